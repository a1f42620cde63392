"""Decoder entry points with the reference's signatures (decoders.py); the sampling runs in
libqecmc's CUDA kernels.  ``*_batch`` variants decode many syndromes per call -- one syndrome
cannot fill a GPU."""
import time

import numpy as np

from . import _lib
from .src import mcmc as _mcmc
from .src.toric_model import Toric_code
from .src.planar_model import Planar_code
from .src.rotated_surface_model import RotSurCode
from .src.xzzx_model import xzzx_code

_seed_counter = [0]


def _next_seed():
    _seed_counter[0] += 1
    return (_mcmc.SEED << 24) ^ (_seed_counter[0] * 0x9E3779B1) ^ (int(time.time_ns()) & 0xFFFFFF if _mcmc.SEED == 0x51ED2020 else 0)


def _as_batch(init_code):
    """Normalise the reference's `init_code` argument (a code object, or a list with one code per
    class, decoders.py:272-279) to (code_class, L, qm [1, ...], per_class)."""
    if type(init_code) == list:
        n_eq = init_code[0].nbr_eq_classes
        assert len(init_code) == n_eq, 'if init_code is a list, it has to contain one code for each class'
        qm = np.stack([np.ascontiguousarray(c.qubit_matrix, dtype=np.uint8).reshape(-1) for c in init_code])[None]
        return init_code[0], np.ascontiguousarray(qm), True
    q = init_code.qubit_matrix
    if not isinstance(q, np.ndarray) or q.dtype != np.uint8:
        raise TypeError("qubit_matrix must be a numpy uint8 array (SURVEY.md Q5)")
    return init_code, np.ascontiguousarray(q).reshape(1, -1), False


def STDC_batch(codes, p_error, p_sampling=None, droplets=10, steps=20000, conv_mult=0, seed=None, device=0,
               return_stats=False):
    """STDC for a list of code objects (same class and size) -> float64 [S, nbr_eq_classes]."""
    p_sampling = p_sampling or p_error
    first = codes[0]
    L = first.system_size
    qm = np.stack([np.ascontiguousarray(c.qubit_matrix, dtype=np.uint8).reshape(-1) for c in codes])
    ctx = _lib.default_context(device)
    out, st = ctx.stdc(first.geometry, _mcmc.fast_path_geometry(first), L, qm, p_error, p_sampling, droplets, int(steps),
                       iters=5, per_class=False, randomize=hasattr(type(first), "apply_stabilizers_uniform") and first.layers == 2,
                       conv_mult=float(conv_mult), seed=_next_seed() if seed is None else seed)
    return (out, st) if return_stats else out


def STDC(init_code, p_error, p_sampling=None, droplets=10, steps=20000, conv_mult=0):
    """decoders.py:268-322.  Returns the class distribution in percent (float64[nbr_eq_classes])."""
    p_sampling = p_sampling or p_error
    code, qm, per_class = _as_batch(init_code)
    if not per_class and code.layers != 2:
        raise TypeError("the fast path only accepts (2, L, L) lattices (SURVEY.md Q5); pass a list of per-class inits")
    ctx = _lib.default_context()
    out, _ = ctx.stdc(code.geometry, _mcmc.fast_path_geometry(code), code.system_size, qm, p_error, p_sampling, droplets,
                      int(steps), iters=5, per_class=per_class, randomize=not per_class, conv_mult=float(conv_mult),
                      seed=_next_seed())
    return out[0]
