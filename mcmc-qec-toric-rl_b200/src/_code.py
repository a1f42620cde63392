"""Shared implementation of the four code objects.

The reference keeps one numba module per code (src/toric_model.py, planar_model.py,
rotated_surface_model.py, xzzx_model.py) with the same duck-typed interface (SURVEY.md
section 1, L1).  Here the geometry of each code is a table of (flat site index, Pauli)
pairs per stabilizer / logical operator, built once per (code, size) with numpy; the
public methods keep the reference's names, argument order and functional style
(every ``apply_*`` returns ``(new_matrix_copy, error_count_delta)`` and leaves
``self.qubit_matrix`` untouched).  These are host utilities around the lattice; the
Metropolis chains and decoders run on the GPU (mcmc.py, decoders.py).
"""
import functools
import random as _pyrandom

import numpy as np

from .. import _lib


def _delta(old, new):
    return int(np.count_nonzero(new) - np.count_nonzero(old))


class CodeBase:
    geometry = None          # _lib.TORIC ...
    nbr_eq_classes = 4
    layers = 1

    def __init__(self, size):
        if self.layers == 1 and (size < 3 or size % 2 == 0):
            raise ValueError("rotated / XZZX codes need an odd size >= 3")
        self.system_size = size
        shape = (2, size, size) if self.layers == 2 else (size, size)
        self.qubit_matrix = np.zeros(shape, dtype=np.uint8)

    # ---- geometry tables ------------------------------------------------
    @classmethod
    @functools.lru_cache(maxsize=None)
    def _stabilizer_table(cls, L):
        """dict (row, col, operator) -> (flat indices, paulis)"""
        raise NotImplementedError

    def _flat(self, *idx):
        L = self.system_size
        if self.layers == 2:
            l, r, c = idx
            return (l * L + r) * L + c
        r, c = idx
        return r * L + c

    def _xor(self, sites, paulis):
        new = self.qubit_matrix.copy()
        flat = new.reshape(-1)
        for s, p in zip(sites, paulis):     # sequential: a site may appear twice (crossing logicals)
            flat[s] ^= p
        return new, _delta(self.qubit_matrix, new)

    # ---- reference interface ---------------------------------------------
    def count_errors(self):
        return int(np.count_nonzero(self.qubit_matrix))

    def chain_lengths(self):
        q = self.qubit_matrix
        return int((q == 1).sum()), int((q == 2).sum()), int((q == 3).sum())

    def count_errors_xyz(self):
        return np.array(self.chain_lengths(), dtype=np.float64)

    def apply_stabilizer(self, row, col, operator):
        sites, paulis = self._stabilizer_table(self.system_size)[(int(row), int(col), int(operator))]
        return self._xor(sites, paulis)

    def apply_random_stabilizer(self):
        table = self._stabilizer_table(self.system_size)
        key = list(table)[_pyrandom.randrange(len(table))]   # every stabilizer is equiprobable
        return self._xor(*table[key])

    def apply_random_logical(self):
        L = self.system_size
        op = _pyrandom.randrange(4)
        return self.apply_logical(op, X_pos=_pyrandom.randrange(L), Z_pos=_pyrandom.randrange(L))

    def define_equivalence_class(self):
        raise NotImplementedError

    def to_class(self, eq):
        """Error chain with the same syndrome in class ``eq`` (decoders.py:556-560 route;
        the reference defines it for the toric code only)."""
        diff = self.define_equivalence_class() ^ int(eq)
        return self.apply_logical(diff)[0]

    def apply_stabilizers_uniform(self, p=0.5):
        raise AttributeError(f"{type(self).__name__} has no apply_stabilizers_uniform (as in the reference)")

    def generate_random_error(self, p_x, p_y=None, p_z=None):
        """i.i.d. Pauli errors.  One argument: depolarizing rate p (X, Y, Z equally likely);
        three arguments: (p_x, p_y, p_z) as in planar_model.py:18-46."""
        if p_y is None:
            p_x = p_y = p_z = p_x / 3.0
        r = np.random.uniform(0, 1, size=self.qubit_matrix.shape)
        q = np.zeros(self.qubit_matrix.shape, dtype=np.uint8)
        q[r < p_z] = 3
        q[(r > p_z) & (r < p_z + p_x)] = 1
        q[(r > p_z + p_x) & (r < p_z + p_x + p_y)] = 2
        self.qubit_matrix = q
        self._clear_unused()

    def generate_zbiased_error(self, p_error, eta):
        pz = p_error * eta / (eta + 1)
        px = p_error / (2 * (eta + 1))
        self.generate_random_error(px, px, pz)

    def generate_alpha_error(self, p_error, alpha):
        """planar_model.py:79-99: p_tilde = p/(1+p); pz_tilde solves x + 2 x**alpha = p_tilde; then
        (p_x, p_y, p_z) = (pz_tilde**alpha, pz_tilde**alpha, pz_tilde) * (1 - p_error), i.i.d. per qubit.
        (The reference's own loop indexes the (2, L, L) planar lattice with two indices and fails for L > 2;
        this draws every qubit of the code.)"""
        from scipy import optimize
        p_tilde = p_error / (1 + p_error)
        pz_tilde = optimize.fsolve(lambda x: x + 2 * x**alpha - p_tilde, 0.5)[0]
        p_xy = pz_tilde**alpha * (1 - p_error)
        self.generate_random_error(p_xy, p_xy, pz_tilde * (1 - p_error))

    def _clear_unused(self):
        pass

    def syndrome(self):
        """Defect indicator per stabilizer, keyed like the stabilizer table."""
        flat = self.qubit_matrix.reshape(-1)
        out = {}
        for key, (sites, paulis) in self._stabilizer_table(self.system_size).items():
            # a stabilizer anticommutes with a qubit error that is non-identity and differs from its own Pauli
            out[key] = int(sum(1 for s, p in zip(sites, paulis) if flat[s] != 0 and flat[s] != p) % 2)
        return out

    syndrom = syndrome
