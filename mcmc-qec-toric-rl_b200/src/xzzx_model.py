"""xzzx_code -- host mirror of the reference's src/xzzx_model.py interface."""
import functools

from .. import _lib
from ._code import CodeBase
from .rotated_surface_model import _half_sites


class xzzx_code(CodeBase):
    geometry = _lib.XZZX
    nbr_eq_classes = 4           # xzzx_model.py:9
    layers = 1

    @classmethod
    @functools.lru_cache(maxsize=None)
    def _stabilizer_table(cls, L):
        t = {}
        for r in range(L - 1):                      # xzzx_model.py:369-381: X Z / Z X on the 2x2 block
            for c in range(L - 1):
                t[(r, c, 1)] = ([r * L + c, (r + 1) * L + c, r * L + c + 1, (r + 1) * L + c + 1], [1, 3, 3, 1])
        half_paulis = {0: [3, 1], 1: [1, 3], 2: [1, 3], 3: [3, 1]}
        for k in range((L - 1) // 2):
            for side in range(4):
                t[(k, side, 3)] = ([a * L + b for a, b in _half_sites(L, k, side)], half_paulis[side])
        return t

    def generate_known_error(self, p_error):
        self.qubit_matrix[0, 1] = 1
        self.qubit_matrix[1, 1] = 1

    def apply_logical(self, operator, X_pos=0, Z_pos=0):
        """operator in {1,2}: X on the anti-diagonal; {2,3}: Z on the diagonal (xzzx_model.py:279-313)."""
        L = self.system_size
        sites, paulis = [], []
        if operator in (1, 2):
            sites += [i * L + (L - 1 - i) for i in range(L)]
            paulis += [1] * L
        if operator in (2, 3):
            sites += [i * L + i for i in range(L)]
            paulis += [3] * L
        return self._xor(sites, paulis)

    def define_equivalence_class(self):
        q = self.qubit_matrix                       # xzzx_model.py:455-486
        L = self.system_size
        x = sum(1 for i in range(L) if q[0, i] == 2 or q[0, i] == (1 if i % 2 == 0 else 3)) % 2
        z = sum(1 for i in range(L) if q[i, 0] == 2 or q[i, 0] == (3 if i % 2 == 0 else 1)) % 2
        return {(0, 0): 0, (1, 0): 1, (1, 1): 2, (0, 1): 3}[(x, z)]
