"""Class-sorted MWPM start states -- mirror of the reference's src/mwpm.py for Planar_code.

The reference builds a defect graph per layer (``MWPM.generate_edges`` :66-133, ``generate_edges_constrained`` :136-229),
writes it to a text file and runs the external ``blossom5`` binary from a hard-coded cluster path (:376-405).  Here the
same graphs are solved in process by the native library (``qecmc_mwpm_planar``, csrc/qecmc_mwpm.cu: a dense primal-dual
blossom algorithm, one syndrome per host thread) -- host code like the reference's, it runs once per syndrome before
the chains start.  Matching weights are the reference's; which of several minimum-weight matchings comes out is the
solver's choice in both.

``MWPM`` on a ``Toric_code`` reads ``code.current_state``, an attribute the reference's ``Toric_code`` does not have
(toric_model.py:13 names it ``defect_matrix``), so only the planar route exists there; the same holds here.
"""
import numpy as np

from .. import _lib
from .planar_model import Planar_code


class MWPM:
    def __init__(self, code):
        assert type(code) is Planar_code, 'code has to be a Planar_code (the toric route of the reference is broken)'
        self.code = code
        self.is_planar = True

    def get_layer(self, layer):                     # mwpm.py:52-63
        return self.code.vertex_defects if layer == 0 else self.code.plaquette_defects

    def _run(self, class_sorted):
        return _lib.mwpm_planar(self.code.system_size, vertex_defects=np.asarray(self.code.vertex_defects)[None],
                                plaquette_defects=np.asarray(self.code.plaquette_defects)[None], class_sorted=class_sorted)

    def solve(self, random_pairing=False):          # mwpm.py:408-415
        if random_pairing:
            raise NotImplementedError("random pairings (mwpm.py:33-50) are not a matching problem; not mirrored")
        return self._run(False)[0][0]

    def generate_classes(self):                     # mwpm.py:417-438; here already in class order
        return list(self._run(True)[0][0])


def class_sorted_mwpm(code):
    """mwpm.py:462-475: four Planar_code objects, entry i a minimum-weight correction of code's syndrome in class i
    (read from code.vertex_defects / code.plaquette_defects: call code.syndrom() first, as the reference requires)."""
    assert type(code) is Planar_code, 'Corrections in different classes can only be generated for planar code'
    out = []
    for chain in MWPM(code).generate_classes():
        c = Planar_code(code.system_size)
        c.qubit_matrix = np.ascontiguousarray(chain)
        out.append(c)
    return out


def class_sorted_mwpm_batch(qubit_matrices, size, threads=0):
    """Batched form: error chains [S][2][L][L] (their syndromes are taken) -> start chains [S][4][2][L][L], class order."""
    return _lib.mwpm_planar(size, qm=qubit_matrices, class_sorted=True, threads=threads)[0]


def regular_mwpm(code):
    """mwpm.py:479-487: the class of the unconstrained minimum-weight correction."""
    sol = type(code)(code.system_size)
    sol.qubit_matrix = np.ascontiguousarray(MWPM(code).solve())
    return sol.define_equivalence_class()


def enhanced_mwpm(code, model="depolarizing", p_xyz=None):
    """mwpm.py:490-517: pick the class whose constrained matching is most likely under the noise model."""
    classes = class_sorted_mwpm(code)
    if model == "depolarizing":
        n = np.array([c.count_errors() for c in classes])
        return np.random.choice(np.where(n == n.min())[0])
    if model == "uncorrelated":
        cnt = [c.count_errors_xyz() for c in classes]
        w = np.array([e[0] + 2 * e[1] + e[2] for e in cnt])
        return np.random.choice(np.where(w == w.min())[0])
    if model == "biased" and p_xyz is not None:
        cnt = [np.asarray(c.count_errors_xyz(), dtype=float) for c in classes]
        rel = (np.asarray(p_xyz) / 3) / (1 - np.asarray(p_xyz))
        prob = np.array([(rel ** (e - cnt[0])).prod() for e in cnt])
        return np.random.choice(np.where(prob == prob.max())[0])
    raise ValueError("unknown model %r" % (model,))
