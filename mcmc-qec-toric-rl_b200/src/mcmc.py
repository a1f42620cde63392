"""Chain / Ladder -- host mirror of the reference's src/mcmc.py; the Metropolis steps run on the GPU.

``FAST_PATH_GEOMETRY`` selects what ``update_chain_fast`` (and the STDC/STRC/single_temp decoders
built on it) proposes with:

* ``"code"`` (default) -- the code's own stabilizers.  This is the reference with the one symbol
  of SURVEY.md Q1 rebound (the patched oracle) and the only setting that decodes toric codes.
* ``"shipped"`` -- planar stabilizers on any (2, L, L) lattice, which is what the unmodified
  reference does (src/mcmc.py:6 imports planar_model._apply_random_stabilizer only).
"""
import copy
import itertools
import random as _pyrandom

import numpy as np

from .. import _lib

FAST_PATH_GEOMETRY = "code"
_chain_ids = itertools.count(1)
SEED = 0x51ED2020


def seed(value):
    """Seed for the per-chain Philox streams of chains / ladders created afterwards."""
    global SEED
    SEED = int(value)


def _new_stream():
    return ((SEED & 0xFFFFFFFFFF) << 24) ^ next(_chain_ids)


def _call_key(stream, calls):
    """64-bit Philox key of call number `calls` of a host-driven ladder / top-rung chain: (stream, calls) mixed by
    splitmix64, so the call counter never runs into the seed bits of `stream` and never wraps out of 64 bits."""
    z = (stream ^ (calls * 0x9E3779B97F4A7C15)) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    return z ^ (z >> 31)


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
        self._stream = _new_stream()
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
        """Chain.update_chain (src/mcmc.py:19-43).  With p_logical != 0 this is a ladder's top rung (random logical
        operators mixed in): it runs as a one-rung ladder block without swaps."""
        if self.p_logical != 0:
            _single_rung_block(self, _lib.LADDER_DEPOLARIZING, self.p, 0.0, iters)
        else:
            self._run(self.code.geometry, iters, _lib.POW_LIBM)


class Chain_xyz:
    """Chain_xyz(p_xyz, code): fast-path chain for general (p_x, p_y, p_z) noise (src/mcmc.py:106-114)."""

    def __init__(self, p_xyz, code):
        self.code = code
        self.p_xyz = np.asarray(p_xyz, dtype=np.float64)
        self.factors = self.p_xyz / (1.0 - self.p_xyz.sum())
        self.qubit_errors = code.count_errors_xyz()
        self._stream = _new_stream()
        self._steps = 0

    def update_chain_fast(self, iters):
        q = np.ascontiguousarray(self.code.qubit_matrix, dtype=np.uint8)
        flat = q.reshape(1, -1).copy()
        _lib.default_context().chain_update_xyz(fast_path_geometry(self.code), self.code.system_size, flat, self.p_xyz, int(iters),
                                                seed=self._stream, stream_offset=self._steps)
        self._steps += int(iters)
        self.code.qubit_matrix = flat.reshape(q.shape)
        self.qubit_errors = self.code.count_errors_xyz()


def _single_rung_block(chain, kind, bottom, param_b, iters):
    """`iters` slow-path steps of one chain = Ladder.step of a one-rung ladder (no swap partner)."""
    q = np.ascontiguousarray(chain.code.qubit_matrix, dtype=np.uint8)
    chain._steps += 1
    out = _lib.default_context().ladder_run(chain.code.geometry, chain.code.system_size, kind, q.reshape(1, -1).copy(), bottom, 1,
                                            1, iters=int(iters), param_b=param_b, p_logical=float(chain.p_logical),
                                            seed=_call_key(chain._stream, chain._steps))
    chain.code.qubit_matrix = out["rung_states"][0, 0].reshape(q.shape)
    return out


class _LadderBase:
    """Shared driver of Ladder / Ladder_alpha / Ladder_biased: the rung states live in numpy between .step() calls;
    every .step(iters) is one device call that runs all rungs and the swap sweep (mcmc.py:94-103)."""
    _kind = _lib.LADDER_DEPOLARIZING

    def _setup(self, init_code, Nc, p_logical, bottom, param_b, ladder, chains):
        self.init_code = init_code
        self.Nc = Nc
        self.p_logical = p_logical
        self.chains = chains
        self.chains[-1].flag = 1
        self.chains[-1].p_logical = p_logical
        self.tops0 = 0
        self._bottom, self._param_b = bottom, param_b
        self._stream = _new_stream()
        self._calls = 0
        self._parts = None

    def _state(self):
        n = self.chains[0].code.qubit_matrix.size
        rs = np.stack([np.ascontiguousarray(c.code.qubit_matrix, dtype=np.uint8).reshape(-1) for c in self.chains])[None]
        flags = np.array([[c.flag for c in self.chains]], np.int32)
        if self._parts is None:  # Chain_alpha.__init__: n_eff from the initial state (mcmc_alpha.py:18-22)
            self._parts = np.array([[[int((c.code.qubit_matrix == 3).sum()), int(((c.code.qubit_matrix == 1) |
                                      (c.code.qubit_matrix == 2)).sum())] for c in self.chains]], np.int32)
        return dict(rung_states=np.ascontiguousarray(rs), flags=flags, tops0=np.array([self.tops0], np.int64),
                    n_eff_parts=np.ascontiguousarray(self._parts))

    def update_ladder(self, iters):
        for chain in self.chains:
            chain.update_chain(iters)

    def step(self, iters):
        code0 = self.chains[0].code
        self._calls += 1
        out = _lib.default_context().ladder_run(code0.geometry, code0.system_size, self._kind, None, self._bottom, self.Nc, 1,
                                                iters=int(iters), param_b=self._param_b, p_logical=float(self.p_logical),
                                                seed=_call_key(self._stream, self._calls), resume=self._state())
        shape = code0.qubit_matrix.shape
        for i, c in enumerate(self.chains):
            c.code.qubit_matrix = out["rung_states"][0, i].reshape(shape).copy()
            c.flag = int(out["flags"][0, i])
            if hasattr(c, "n_eff"):
                c.n_eff = float(out["n_eff"][0, i])
        self._parts = out["n_eff_parts"]
        self.tops0 = int(out["tops0"][0])


class Ladder(_LadderBase):
    """Ladder(p_bottom, init_code, Nc, p_logical=0): src/mcmc.py:49-103."""

    def __init__(self, p_bottom, init_code, Nc, p_logical=0):
        self.p_bottom = p_bottom
        p_ladder = np.linspace(p_bottom, 0.75, Nc)
        self.p_ladder = p_ladder
        self.p_diff = (p_ladder[:-1] * (1 - p_ladder[1:])) / (p_ladder[1:] * (1 - p_ladder[:-1]))
        self._setup(init_code, Nc, p_logical, p_bottom, 0.0, p_ladder, [Chain(p, copy.deepcopy(init_code)) for p in p_ladder])

    def r_flip(self, ind_lo):
        """Swap decision for rungs (ind_lo, ind_lo + 1) from the rungs' current lengths (src/mcmc.py:86-92, 144-149).
        Host-side (one decision is not GPU work); Ladder.step takes its own decisions on the device."""
        ne_lo = self.chains[ind_lo].code.count_errors()
        ne_hi = self.chains[ind_lo + 1].code.count_errors()
        if ne_hi < ne_lo:
            return True
        return _pyrandom.random() < float(self.p_diff[ind_lo]) ** (ne_hi - ne_lo)


class MCMCDataReader:
    """Cursor over a data set written by generate_data.generate (or by the reference): a pickled DataFrame indexed by
    (data_nr, type).  Same constructor and methods as the reference's reader (src/mcmc.py:118-141); a missing or
    unreadable file raises instead of leaving a half-built object behind."""

    def __init__(self, file_path, size):
        import pandas as pd
        self._path, self._size, self._pos = file_path, size, 0
        try:
            self._frame = pd.read_pickle(file_path)
        except (OSError, ValueError, EOFError) as exc:
            raise FileNotFoundError(f"MCMCDataReader: cannot read {file_path!r}: {exc}") from exc
        last_nr = self._frame.index[-1][0]
        self._count = int(last_nr) + 1

    def full(self):
        """Every cell of the frame as one flat object array (params row first)."""
        return self._frame.to_numpy().ravel()

    def has_next(self):
        return self._pos < self._count

    def current_index(self):
        return self._pos

    def get_capacity(self):
        return self._count
