"""Decoder entry points with the reference's signatures (decoders.py); the sampling runs in libqecmc's CUDA
kernels.  ``*_batch`` variants decode many syndromes per call -- one syndrome cannot fill a GPU -- and return
one row per syndrome.

init_code is a code object or a list with one code object per equivalence class (decoders.py:272-279);
``p_sampling or p_error`` means p_sampling = 0 falls back to p_error, as in the reference (decoders.py:270).
"""
import time

import numpy as np

from . import _lib
from .src import mcmc as _mcmc
from .src.toric_model import Toric_code            # noqa: F401  (re-exported like the reference's star imports)
from .src.planar_model import Planar_code          # noqa: F401
from .src.rotated_surface_model import RotSurCode  # noqa: F401
from .src.xzzx_model import xzzx_code              # noqa: F401
from .src.mcmc import Chain, Ladder, Chain_xyz     # noqa: F401
from .src.mcmc_alpha import Chain_alpha, Ladder_alpha      # noqa: F401
from .src.mcmc_biased import Chain_biased, Ladder_biased   # noqa: F401

_seed_counter = [0]


def _next_seed(seed=None):
    if seed is not None:
        return int(seed)
    _seed_counter[0] += 1
    salt = (int(time.time_ns()) & 0xFFFFFF) if _mcmc.SEED == 0x51ED2020 else 0   # unseeded like the reference unless mcmc.seed() was called
    return ((_mcmc.SEED & 0xFFFFFFFF) << 24) ^ (_seed_counter[0] * 0x9E3779B1) ^ salt


def _u8(code):
    q = code.qubit_matrix
    if not isinstance(q, np.ndarray) or q.dtype != np.uint8:
        raise TypeError("qubit_matrix must be a numpy uint8 array (the reference's njit signatures, SURVEY.md Q5)")
    return np.ascontiguousarray(q).reshape(-1)


def _batch(init_codes):
    """list of init_code arguments (code objects, or per-class lists) -> (first code, qm, per_class)."""
    first = init_codes[0]
    if type(first) == list:
        n_eq = first[0].nbr_eq_classes
        for ic in init_codes:
            assert len(ic) == n_eq, 'if init_code is a list, it has to contain one code for each class'
        qm = np.stack([np.stack([_u8(c) for c in ic]) for ic in init_codes])
        return first[0], np.ascontiguousarray(qm), True
    return first, np.ascontiguousarray(np.stack([_u8(c) for c in init_codes])), False


def _fast_path_check(code, per_class):
    if code.layers != 2:
        raise TypeError("update_chain_fast only accepts (2, L, L) lattices (numba signature uint8[:,:,:], SURVEY.md Q5)")
    # a single planar code is accepted (classes reached via apply_logical(class ^ eq), decoders.py:556-560); the
    # reference raises AttributeError there because only Toric_code has to_class


# ------------------------------------------------------------------------------------------------ STDC / STRC
def STDC_batch(init_codes, p_error, p_sampling=None, droplets=10, steps=20000, conv_mult=0, seed=None, device=0,
               return_stats=False):
    """STDC (decoders.py:268-322) for a list of init_code arguments -> float64 [S, nbr_eq_classes] (percent).
    conv_mult != 0 is the early stop of decoders.py:257-263 ("new" = new to the droplet: a set per chain on the device)."""
    p_sampling = p_sampling or p_error
    code, qm, per_class = _batch(init_codes)
    _fast_path_check(code, per_class)
    out, st = _lib.default_context(device).stdc(
        code.geometry, _mcmc.fast_path_geometry(code), code.system_size, qm, p_error, p_sampling, int(droplets), int(steps),
        iters=5, per_class=per_class, randomize=not per_class, conv_mult=float(conv_mult), seed=_next_seed(seed))
    return (out, st) if return_stats else out


def STDC(init_code, p_error, p_sampling=None, droplets=10, steps=20000, conv_mult=0):
    return STDC_batch([init_code], p_error, p_sampling, droplets, steps, conv_mult)[0]


def STRC_batch(init_codes, p_error, p_sampling=None, droplets=10, steps=20000, conv_mult=0, seed=None, device=0,
               return_stats=False):
    """STRC (decoders.py:835-949) for a list of init_code arguments -> float64 [S, nbr_eq_classes] (percent)."""
    p_sampling = p_sampling or p_error
    code, qm, per_class = _batch(init_codes)
    _fast_path_check(code, per_class)
    out, st = _lib.default_context(device).strc(
        code.geometry, _mcmc.fast_path_geometry(code), code.system_size, qm, p_error, p_sampling, int(droplets), int(steps),
        iters=5, per_class=per_class, randomize=not per_class, conv_mult=float(conv_mult), seed=_next_seed(seed))
    return (out, st) if return_stats else out


def STRC(init_code, p_error, p_sampling=None, droplets=10, steps=20000, conv_mult=0):
    return STRC_batch([init_code], p_error, p_sampling, droplets, steps, conv_mult)[0]


def single_temp_batch(init_codes, p, max_iters, seed=None, device=0):
    """single_temp (decoders.py:108-135): mean chain length per class -> float64 [S, nbr_eq_classes]."""
    code, qm, per_class = _batch(init_codes)
    _fast_path_check(code, per_class)
    out, _ = _lib.default_context(device).single_temp(code.geometry, _mcmc.fast_path_geometry(code), code.system_size, qm, p,
                                                      int(max_iters), iters=5, per_class=per_class, seed=_next_seed(seed))
    return out


def single_temp(init_code, p, max_iters):
    return single_temp_batch([init_code], p, max_iters)[0]


# ------------------------------------------------------------------------------------------------ PTEQ / PTDC
def conv_crit_error_based_PT(nbr_errors_bottom_chain, since_burn, tops_accepted, SEQ, eps):
    """The reference's convergence test on a host array (decoders.py:91-104): the history of the bottom rung's length has
    since_burn + 1 valid entries; its second and fourth quarters must agree within eps.  -> (agree, agree and
    tops_accepted >= SEQ).  The device evaluates the same windows as running integer sums inside the tempering kernel; this
    helper is for callers that keep their own history."""
    n = int(since_burn) + 1
    hist = np.asarray(nbr_errors_bottom_chain)
    with np.errstate(invalid="ignore", divide="ignore"):
        gap = abs(np.mean(hist[n // 4: n // 2]) - np.mean(hist[3 * n // 4: n])) if n >= 2 else float("nan")
    agree = bool(gap < eps)
    return agree, bool(agree and tops_accepted >= SEQ)


def PTEQ_batch(init_codes, p, Nc=None, SEQ=2, TOPS=10, tops_burn=2, eps=0.1, steps=50000000, iters=10,
               conv_criteria='error_based', seed=None, device=0, return_info=False, _kind=_lib.LADDER_DEPOLARIZING,
               _param_b=0.0):
    """PTEQ (decoders.py:25-89) for a list of code objects -> uint8 [S, nbr_eq_classes] (truncated percent).

    `steps` caps the Ladder.step calls per syndrome.  With conv_criteria='error_based' the device keeps 2 bytes of history
    per step (4 for alpha ladders) for every ladder in flight -- a finished ladder hands its place to the next one of the
    batch -- so the reference's default cap of 5e7 runs (with fewer ladders resident), it just takes as long as its slowest
    ladder."""
    code, qm, per_class = _batch(init_codes)
    if per_class:
        raise TypeError("PTEQ takes a single code object per syndrome")
    Nc = Nc or code.system_size
    if tops_burn >= TOPS:
        print('tops_burn has to be smaller than TOPS')
    pct, info = _lib.default_context(device).pteq(
        code.geometry, code.system_size, _kind, qm, p, Nc=Nc, param_b=_param_b, SEQ=SEQ, TOPS=TOPS, tops_burn=tops_burn,
        eps=eps, steps=int(steps), iters=int(iters), conv=(conv_criteria == 'error_based'), p_logical=0.5, seed=_next_seed(seed))
    if conv_criteria == 'error_based' and not info["converged"].all():
        print('\n\nWARNING: PTEQ has not converged.\n\n')
    return (pct, info) if return_info else pct


def PTEQ(init_code, p, Nc=None, SEQ=2, TOPS=10, tops_burn=2, eps=0.1, steps=50000000, iters=10, conv_criteria='error_based'):
    return PTEQ_batch([init_code], p, Nc, SEQ, TOPS, tops_burn, eps, steps, iters, conv_criteria)[0]


def PTDC_batch(init_codes, p_error, p_sampling=None, droplets=4, Nc=None, steps=20000, conv_mult=0, seed=None, device=0,
               return_float=False):
    """PTDC (decoders.py:168-233) -> uint8 [S, nbr_eq_classes] (truncated percent)."""
    p_sampling = p_sampling or p_error
    code, qm, per_class = _batch(init_codes)
    Nc = Nc or code.system_size
    steps = int(steps) // Nc                       # decoders.py:196
    out, _ = _lib.default_context(device).ptdc(code.geometry, code.system_size, qm, p_error, p_sampling, int(droplets), Nc,
                                               steps, iters=10, per_class=per_class, seed=_next_seed(seed),
                                               conv_mult=float(conv_mult))
    return out if return_float else out.astype(np.uint8)


def PTDC(init_code, p_error, p_sampling=None, droplets=4, Nc=None, steps=20000, conv_mult=0):
    return PTDC_batch([init_code], p_error, p_sampling, droplets, Nc, steps, conv_mult)[0]


def PTRC_batch(init_codes, p_error, p_sampling=None, droplets=4, Nc=None, steps=20000, conv_mult=2.0, seed=None, device=0,
               return_float=False):
    """PTRC (decoders.py:638-742) -> uint8 [S, nbr_eq_classes] (truncated percent).  conv_mult is accepted and, as in
    the reference (whose early-stop lines are commented out, decoders.py:628-631), has no effect.  A single code
    object is accepted (classes reached with to_class); the reference raises TypeError there."""
    p_sampling = p_sampling or p_error
    code, qm, per_class = _batch(init_codes)
    Nc = Nc or code.system_size
    steps = int(steps) // Nc                       # decoders.py:666
    out, _ = _lib.default_context(device).ptrc(code.geometry, code.system_size, qm, p_error, p_sampling, int(droplets), Nc,
                                               steps, iters=10, per_class=per_class, seed=_next_seed(seed))
    return out if return_float else out.astype(np.uint8)


def PTRC(init_code, p_error, p_sampling=None, droplets=4, Nc=None, steps=20000, conv_mult=2.0):
    return PTRC_batch([init_code], p_error, p_sampling, droplets, Nc, steps, conv_mult)[0]


# ------------------------------------------------------------------------------------------------ EWD-style
def STDC_Nall_n_alpha_batch(init_codes, pz_tilde_sampling=None, alpha=1, pz_tilde=0.1, steps=20000, seed=None, device=0):
    """STDC_Nall_n_alpha (decoders.py:537-581) -> float64 [S, nbr_eq_classes] (percent).  A single toric code is
    accepted here (classes reached with Toric_code.to_class); the reference raises on it."""
    code, qm, per_class = _batch(init_codes)
    out, _, _ = _lib.default_context(device).stdc_alpha(code.geometry, code.system_size, qm, pz_tilde_sampling, alpha, pz_tilde,
                                                        int(steps), iters=5, per_class=per_class, seed=_next_seed(seed))
    return out


def STDC_Nall_n_alpha(init_code, pz_tilde_sampling=None, alpha=1, pz_tilde=0.1, steps=20000):
    return STDC_Nall_n_alpha_batch([init_code], pz_tilde_sampling, alpha, pz_tilde, steps)[0]


# ------------------------------------------------------------------------------------------------ general noise
def STDC_general_noise_shortest_batch(init_codes, p_xyz, p_sampling=None, droplets=10, steps=20000, seed=None, device=0):
    """STDC_general_noise_shortest (decoders.py:435-508) over a batch -> (eqdistr, eqdistr_shortest), float64 [S, n_eq]."""
    p_xyz = np.asarray(p_xyz, dtype=np.float64)
    if p_sampling is None:
        p_sampling = float(p_xyz.sum())            # decoders.py:438-439: a scalar, so the sampling chain is a plain Chain
    code, qm, per_class = _batch(init_codes)
    _fast_path_check(code, per_class)
    out, out_s, _, _ = _lib.default_context(device).stdc_general_noise(
        code.geometry, _mcmc.fast_path_geometry(code), code.system_size, qm, p_xyz,
        np.asarray(p_sampling, dtype=np.float64) if type(p_sampling) == np.ndarray else float(p_sampling), int(droplets), int(steps),
        iters=5, per_class=per_class, seed=_next_seed(seed))
    return out, out_s


def STDC_general_noise_shortest(init_code, p_xyz, p_sampling=None, droplets=10, steps=20000):
    a, b = STDC_general_noise_shortest_batch([init_code], p_xyz, p_sampling, droplets, steps)
    return a[0], b[0]


def STDC_general_noise(init_code, p_xyz, p_sampling=None, droplets=10, steps=20000, shortest_only=False):
    """decoders.py:345-432; shortest_only keeps only the chains whose weighted length is np.isclose to the minimum."""
    a, b = STDC_general_noise_shortest_batch([init_code], p_xyz, p_sampling, droplets, steps)
    return b[0] if shortest_only else a[0]
