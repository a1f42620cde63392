"""Chain / Ladder -- host mirror of the reference's src/mcmc.py; the Metropolis steps run on the GPU.

``FAST_PATH_GEOMETRY`` selects what ``update_chain_fast`` (and the STDC/STRC/single_temp decoders
built on it) proposes with:

* ``"code"`` (default) -- the code's own stabilizers.  This is the reference with the one symbol
  of SURVEY.md Q1 rebound (the patched oracle) and the only setting that decodes toric codes.
* ``"shipped"`` -- planar stabilizers on any (2, L, L) lattice, which is what the unmodified
  reference does (src/mcmc.py:6 imports planar_model._apply_random_stabilizer only).
"""
import itertools

import numpy as np

from .. import _lib

FAST_PATH_GEOMETRY = "code"
_chain_ids = itertools.count(1)
SEED = 0x51ED2020


def seed(value):
    """Seed for the per-chain Philox streams of chains created afterwards."""
    global SEED
    SEED = int(value)


def fast_path_geometry(code):
    if FAST_PATH_GEOMETRY == "shipped":
        if code.layers != 2:
            raise TypeError("the fast path only accepts (2, L, L) lattices (numba signature uint8[:,:,:], SURVEY.md Q5)")
        return _lib.PLANAR
    return code.geometry


class Chain:
    """Chain(p, code): one Metropolis chain at error rate p (src/mcmc.py:10-46)."""

    def __init__(self, p, code):
        self.code = code
        self.p = p
        self.p_logical = 0
        self.flag = 0
        self.factor = ((self.p / 3.0) / (1.0 - self.p))
        self._stream = (SEED << 20) ^ next(_chain_ids)
        self._steps = 0

    def _run(self, geom, iters, pow_kind):
        q = np.ascontiguousarray(self.code.qubit_matrix, dtype=np.uint8)
        flat = q.reshape(1, -1).copy()
        _lib.default_context().chain_update(geom, self.code.system_size, flat, self.p, int(iters), seed=self._stream,
                                            stream_offset=self._steps, pow_kind=pow_kind)
        self._steps += int(iters)
        self.code.qubit_matrix = flat.reshape(q.shape)

    def update_chain_fast(self, iters):
        """_update_chain_fast (src/mcmc.py:152-160): `iters` Metropolis steps on the GPU."""
        self._run(fast_path_geometry(self.code), iters, _lib.POW_NUMBA)

    def update_chain(self, iters):
        """Chain.update_chain (src/mcmc.py:19-43).  The p_logical != 0 branch belongs to the top rung of a
        Ladder and runs inside the ladder kernels; a free-standing chain has p_logical == 0."""
        if self.p_logical != 0:
            raise NotImplementedError("a chain with p_logical != 0 is a ladder's top rung: use Ladder.step / PTEQ")
        self._run(self.code.geometry, iters, _lib.POW_LIBM)
