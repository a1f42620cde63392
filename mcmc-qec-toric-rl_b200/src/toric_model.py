"""Toric_code -- host mirror of the reference's src/toric_model.py interface."""
import functools

import numpy as np

from .. import _lib
from ._code import CodeBase, _delta


class Toric_code(CodeBase):
    geometry = _lib.TORIC
    nbr_eq_classes = 16          # toric_model.py:8
    layers = 2

    @classmethod
    @functools.lru_cache(maxsize=None)
    def _stabilizer_table(cls, L):
        f = lambda l, r, c: (l * L + r % L) * L + c % L
        t = {}
        for r in range(L):
            for c in range(L):
                t[(r, c, 1)] = ([f(1, r, c), f(1, r, c - 1), f(0, r, c), f(0, r - 1, c)], [1] * 4)   # toric_model.py:262-265
                t[(r, c, 3)] = ([f(1, r, c), f(0, r, c), f(0, r, c + 1), f(1, r + 1, c)], [3] * 4)   # toric_model.py:267-270
        return t

    def generate_random_error(self, p_error):
        """Depolarizing errors of rate p_error (distribution of toric_model.py:15-24)."""
        CodeBase.generate_random_error(self, p_error)

    def generate_n_random_errors(self, n):
        flat = np.zeros(self.qubit_matrix.size, dtype=np.uint8)
        flat[:n] = np.random.randint(3, size=n) + 1
        np.random.shuffle(flat)
        self.qubit_matrix = flat.reshape(self.qubit_matrix.shape)

    def apply_logical(self, operator, layer=0, X_pos=0, Z_pos=0):
        """operator in {1,2}: X along row X_pos of `layer`; {2,3}: Z along column Z_pos
        (layer 1 addressed transposed), toric_model.py:179-225.  Unlike the reference's
        wrapper (which drops `layer`) the layer argument is honoured."""
        L = self.system_size
        sites, paulis = [], []
        if operator in (1, 2):
            sites += [self._flat(0, X_pos, i) if layer == 0 else self._flat(1, i, X_pos) for i in range(L)]
            paulis += [1] * L
        if operator in (2, 3):
            sites += [self._flat(0, i, Z_pos) if layer == 0 else self._flat(1, Z_pos, i) for i in range(L)]
            paulis += [3] * L
        return self._xor(sites, paulis)

    def apply_random_logical(self):
        import random
        L = self.system_size
        new, total = self.qubit_matrix, 0
        keep = self.qubit_matrix
        for layer in (0, 1):                      # one operator per layer (toric_model.py:228-253)
            op = random.randrange(4)
            self.qubit_matrix = new
            new, d = self.apply_logical(op, layer, random.randrange(L), random.randrange(L))
            total += d
        self.qubit_matrix = keep
        return new, total

    def define_equivalence_class(self):
        q = self.qubit_matrix                      # toric_model.py:317-351
        bits = [((q[l] == 1) | (q[l] == 2)).sum() % 2 if k == 0 else ((q[l] == 3) | (q[l] == 2)).sum() % 2
                for l in (0, 1) for k in (0, 1)]
        return int(bits[0] + 2 * bits[1] + 4 * bits[2] + 8 * bits[3])

    def to_class(self, eq):
        diff = int(eq) ^ self.define_equivalence_class()    # toric_model.py:354-377
        ops = diff ^ ((diff & 0b1010) >> 1)
        keep = self.qubit_matrix
        self.qubit_matrix = self.apply_logical(ops & 3, 0)[0]
        out = self.apply_logical(ops >> 2, 1)[0]
        self.qubit_matrix = keep
        return out

    def apply_stabilizers_uniform(self, p=0.5):
        L = self.system_size                       # toric_model.py:299-314: index 0 -> operator 3, 1 -> operator 1
        hits = np.random.rand(2, L, L) < p
        new = self.qubit_matrix.copy()
        flat = new.reshape(-1)
        table = self._stabilizer_table(L)
        for o, r, c in zip(*np.nonzero(hits)):
            sites, paulis = table[(int(r), int(c), 3 if o == 0 else 1)]
            for s, pl in zip(sites, paulis):
                flat[s] ^= pl
        return new
