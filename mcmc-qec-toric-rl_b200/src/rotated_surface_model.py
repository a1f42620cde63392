"""RotSurCode -- host mirror of the reference's src/rotated_surface_model.py interface."""
import functools

from .. import _lib
from ._code import CodeBase


def _half_sites(L, k, side):
    """boundary two-qubit stabilizers (rotated_surface_model.py:366-378 / xzzx_model.py:382-434)"""
    return {0: [(0, 2 * k + 1), (0, 2 * k + 2)], 1: [(2 * k + 1, L - 1), (2 * k + 2, L - 1)],
            2: [(L - 1, 2 * k), (L - 1, 2 * k + 1)], 3: [(2 * k, 0), (2 * k + 1, 0)]}[side]


class RotSurCode(CodeBase):
    geometry = _lib.ROTATED
    nbr_eq_classes = 4           # rotated_surface_model.py:9
    layers = 1

    def generate_known_error(self, p_error, eta):
        """rotated_surface_model.py:79-82: two fixed X errors (the arguments are ignored there as well)."""
        self.qubit_matrix[2, 2] = 1
        self.qubit_matrix[1, 0] = 1
        self.syndrome()

    @classmethod
    @functools.lru_cache(maxsize=None)
    def _stabilizer_table(cls, L):
        t = {}
        for r in range(L - 1):
            for c in range(L - 1):
                p = 1 if (r + c) % 2 == 0 else 3
                t[(r, c, 1)] = ([r * L + c, r * L + c + 1, (r + 1) * L + c, (r + 1) * L + c + 1], [p] * 4)
        for k in range((L - 1) // 2):
            for side in range(4):
                p = 1 if side in (0, 2) else 3
                t[(k, side, 3)] = ([a * L + b for a, b in _half_sites(L, k, side)], [p, p])
        return t

    def apply_logical(self, operator, X_pos=0, Z_pos=0):
        """operator in {1,3}: X on column X_pos; {2,3}: Z on row Z_pos (rotated_surface_model.py:251-282)."""
        L = self.system_size
        sites, paulis = [], []
        if operator in (1, 3):
            sites += [i * L + X_pos for i in range(L)]
            paulis += [1] * L
        if operator in (2, 3):
            sites += [Z_pos * L + i for i in range(L)]
            paulis += [3] * L
        return self._xor(sites, paulis)

    def define_equivalence_class(self):
        q = self.qubit_matrix                       # rotated_surface_model.py:411-420
        x = int(((q[0, :] == 1) | (q[0, :] == 2)).sum() % 2)
        z = int(((q[:, 0] == 3) | (q[:, 0] == 2)).sum() % 2)
        return x + 2 * z
