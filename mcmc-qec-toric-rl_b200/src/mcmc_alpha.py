"""Chain_alpha / Ladder_alpha -- host mirror of the reference's src/mcmc_alpha.py (alpha-weighted noise:
px = py = pz_tilde**alpha (1-p), pz = pz_tilde (1-p)).  The accept-ratio denominator is frozen for the
duration of one update_chain call, as in the reference."""
import copy

import numpy as np

from .. import _lib
import random as _pyrandom

from .mcmc import _LadderBase, _new_stream, _single_rung_block
from .mcmc import MCMCDataReader  # noqa: F401  (the reference repeats the reader in this module)


class Chain_alpha:
    def __init__(self, pz_tilde, alpha, code):
        self.code = code
        self.pz_tilde = pz_tilde
        self.alpha = alpha
        self.p_logical = 0
        self.flag = 0
        q = self.code.qubit_matrix
        self.n_eff = int((q == 3).sum()) + self.alpha * (int((q == 1).sum()) + int((q == 2).sum()))   # mcmc_alpha.py:18-22
        self._stream = _new_stream()
        self._steps = 0

    def update_chain(self, iters):
        """mcmc_alpha.py:27-70, `iters` steps on the GPU."""
        out = _single_rung_block(self, _lib.LADDER_ALPHA, self.pz_tilde, self.alpha, iters)
        # n_eff is only rewritten on an accepted move (mcmc_alpha.py:56,70); the device reports the rung-owned
        # value, which for a chain that started from its own state is the same thing
        self.n_eff = float(out["n_eff"][0, 0])

    def update_chain_fast(self, iters):
        """mcmc_alpha.py:73-74 reads self.factor, which Chain_alpha never sets (its assignment is commented out,
        mcmc_alpha.py:21): the reference raises AttributeError here, and so does this mirror."""
        raise AttributeError("'Chain_alpha' object has no attribute 'factor'")


class Ladder_alpha(_LadderBase):
    """Ladder_alpha(pz_tilde_bottom, init_code, alpha, Nc, p_logical=0): src/mcmc_alpha.py:77-137."""
    _kind = _lib.LADDER_ALPHA

    def __init__(self, pz_tilde_bottom, init_code, alpha, Nc, p_logical=0):
        self.alpha = alpha
        self.pz_tilde_bottom = pz_tilde_bottom
        lad = np.linspace(pz_tilde_bottom, 1, Nc)
        self.pz_tilde_ladder = lad
        with np.errstate(divide="ignore", invalid="ignore"):
            self.pz_tilde_diff = (lad[:-1] * (1 - lad[1:])) / (lad[1:] * (1 - lad[:-1]))
        self._setup(init_code, Nc, p_logical, pz_tilde_bottom, alpha, lad,
                    [Chain_alpha(pz, alpha, copy.deepcopy(init_code)) for pz in lad])

    def r_flip(self, ind_lo):
        """mcmc_alpha.py:117-123: always draws; the exponent is the difference of the rung-owned n_eff values."""
        lo, hi = self.chains[ind_lo], self.chains[ind_lo + 1]
        return _pyrandom.random() < (lo.pz_tilde / hi.pz_tilde) ** (hi.n_eff - lo.n_eff)
