"""Chain_biased / Ladder_biased -- host mirror of the reference's src/mcmc_biased.py (Z-biased noise:
pz = p eta/(eta+1), px = py = p/(2(eta+1))).  Frozen accept-ratio denominator per call."""
import copy

import numpy as np

from .. import _lib
import random as _pyrandom

from .mcmc import Chain, _LadderBase, _new_stream, _single_rung_block, fast_path_geometry
from .mcmc import MCMCDataReader  # noqa: F401  (the reference repeats the reader in this module)


class Chain_biased:
    def __init__(self, p, eta, code):
        self.code = code
        self.p = p
        self.eta = eta
        self.p_logical = 0
        self.flag = 0
        self.factor = ((self.p / 3.0) / (1.0 - self.p))   # mcmc_biased.py:17: the fast path samples depolarizing weights
        self._stream = _new_stream()
        self._steps = 0

    def update_chain(self, iters):
        """mcmc_biased.py:21-59, `iters` steps on the GPU."""
        _single_rung_block(self, _lib.LADDER_BIASED, self.p, self.eta, iters)

    def update_chain_fast(self, iters):
        """mcmc_biased.py:62-63: _update_chain_fast with the depolarizing `factor` (not the biased weights)."""
        Chain._run(self, fast_path_geometry(self.code), iters, _lib.POW_NUMBA)


class Ladder_biased(_LadderBase):
    """Ladder_biased(p_bottom, init_code, eta, Nc, p_logical=0): src/mcmc_biased.py:66-124."""
    _kind = _lib.LADDER_BIASED

    def __init__(self, p_bottom, init_code, eta, Nc, p_logical=0):
        self.eta = eta
        p_top = (eta + 1) / (2 * eta + 1)
        p_ladder = np.linspace(p_bottom, p_top, Nc)
        self.p_ladder = p_ladder
        self.p_diff = (p_ladder[:-1] * (1 - p_ladder[1:])) / (p_ladder[1:] * (1 - p_ladder[:-1]))
        self._setup(init_code, Nc, p_logical, p_bottom, eta, p_ladder,
                    [Chain_biased(p, eta, copy.deepcopy(init_code)) for p in p_ladder])

    def r_flip(self, ind_lo):
        """mcmc_biased.py:107-113, 154-156: always draws (no shortcut for a lighter upper replica)."""
        ne_lo = self.chains[ind_lo].code.count_errors()
        ne_hi = self.chains[ind_lo + 1].code.count_errors()
        return _pyrandom.random() < float(self.p_diff[ind_lo]) ** (ne_hi - ne_lo)
