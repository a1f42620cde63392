"""Planar_code -- host mirror of the reference's src/planar_model.py interface."""
import functools

import numpy as np

from .. import _lib
from ._code import CodeBase


class Planar_code(CodeBase):
    geometry = _lib.PLANAR
    nbr_eq_classes = 4           # planar_model.py:10
    layers = 2

    @classmethod
    @functools.lru_cache(maxsize=None)
    def _stabilizer_table(cls, L):
        f = lambda l, r, c: (l * L + r) * L + c
        t = {}
        for r in range(L - 1):                      # planar_model.py:300-313
            for c in range(L):
                s = [f(0, r, c), f(0, r + 1, c)]
                if c < L - 1:
                    s.append(f(1, r, c))
                if c > 0:
                    s.append(f(1, r, c - 1))
                t[(r, c, 1)] = (s, [1] * len(s))
        for r in range(L):                          # planar_model.py:315-328
            for c in range(L - 1):
                s = [f(0, r, c), f(0, r, c + 1)]
                if r < L - 1:
                    s.append(f(1, r, c))
                if r > 0:
                    s.append(f(1, r - 1, c))
                t[(r, c, 3)] = (s, [3] * len(s))
        return t

    def __init__(self, size):
        super().__init__(size)
        self.plaquette_defects = np.zeros((size, size - 1), dtype=bool)    # planar_model.py:15-16
        self.vertex_defects = np.zeros((size - 1, size), dtype=bool)

    def syndrom(self):
        """planar_model.py:134-153: vertex defects [L-1][L] from Y / Z errors, plaquette defects [L][L-1] from X / Y
        errors, kept on the object for src/mwpm.py; returns the per-stabilizer dict of the other codes as well."""
        q = self.qubit_matrix
        yz = (q == 2) | (q == 3)
        self.vertex_defects = (yz[0, 1:, :] ^ yz[0, :-1, :]) ^ (yz[1, :-1, :] ^ np.roll(yz[1, :-1, :], 1, axis=1))
        xy = (q == 1) | (q == 2)
        self.plaquette_defects = (xy[0, :, 1:] ^ xy[0, :, :-1]) ^ (xy[1, :, :-1] ^ np.roll(xy[1, :, :-1], 1, axis=0))
        return super().syndrome()

    syndrome = syndrom

    def _clear_unused(self):
        self.qubit_matrix[1, -1, :] = 0             # layer 1 lives in [0:L-1, 0:L-1] (planar_model.py:36-37)
        self.qubit_matrix[1, :, -1] = 0

    def generate_general_noise_error(self, p_xyz):
        self.generate_random_error(p_xyz[0], p_xyz[1], p_xyz[2])

    def generate_biased_error(self, p_error, eta):
        self.generate_zbiased_error(p_error, eta)

    def apply_logical(self, operator, X_pos=0, Z_pos=0):
        """operator in {1,3}: X on [0,X_pos,:]; {2,3}: Z on [0,:,Z_pos] (planar_model.py:234-268)."""
        L = self.system_size
        sites, paulis = [], []
        if operator in (1, 3):
            sites += [self._flat(0, X_pos, i) for i in range(L)]
            paulis += [1] * L
        if operator in (2, 3):
            sites += [self._flat(0, i, Z_pos) for i in range(L)]
            paulis += [3] * L
        return self._xor(sites, paulis)

    def define_equivalence_class(self):
        q = self.qubit_matrix                       # planar_model.py:379-390
        x = int(((q[0, :, 0] == 1) | (q[0, :, 0] == 2)).sum() % 2)
        z = int(((q[0, 0, :] == 3) | (q[0, 0, :] == 2)).sum() % 2)
        return x + 2 * z

    def apply_stabilizers_uniform(self, p=0.5):
        L = self.system_size                        # planar_model.py:355-376
        hits = np.random.rand(2, L, L) < p
        hits[1, L - 1, :] = False
        hits[0, :, L - 1] = False
        new = self.qubit_matrix.copy()
        flat = new.reshape(-1)
        table = self._stabilizer_table(L)
        for o, r, c in zip(*np.nonzero(hits)):
            sites, paulis = table[(int(r), int(c), 3 if o == 0 else 1)]
            for s, pl in zip(sites, paulis):
                flat[s] ^= pl
        return new
