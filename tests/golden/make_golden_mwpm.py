"""Golden vectors for the class-sorted MWPM initialiser -- run in the build container, where /root/reference exists.

The reference's src/mwpm.py builds its defect graphs in Python and hands them to the external blossom5 binary
(mwpm.py:376-405, hard-coded cluster path), which is absent.  Here MWPM.generate_MWPM -- that one call -- is replaced by
networkx.min_weight_matching on the very edge list the reference built; everything else (generate_edges,
generate_edges_constrained, solve_layer, eliminate_*, generate_classes, class_sorted_mwpm, regular_mwpm) is the
unmodified reference.  A minimum-weight matching is not unique, so what is recorded is what every exact solver must
agree on: per (layer, parity) the weight of the constrained matching (= Pauli count of solve_layer's correction), the
weight of the free matching per layer, and the defects; plus one solver's chains for reference.

    python tests/golden/make_golden_mwpm.py      -> tests/golden/mwpm_planar.npz
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from oracle import refshim  # noqa: E402

ref = refshim.load("shipped")
import networkx as nx  # noqa: E402
import src.mwpm as mwpm  # noqa: E402


def nx_generate_MWPM(self, layer, edges, nbr_nodes):
    g = nx.Graph()
    g.add_nodes_from(range(nbr_nodes))
    for a, b, w in edges:
        g.add_edge(int(a), int(b), weight=int(w))
    m = nx.min_weight_matching(g)
    assert len(m) * 2 == nbr_nodes, "no perfect matching"
    out = np.array([[min(a, b), max(a, b)] for a, b in m], dtype=int).reshape(-1, 2)
    return out[np.argsort(out[:, 0])]


mwpm.MWPM.generate_MWPM = nx_generate_MWPM
Planar = ref.planar_model.Planar_code

cases = [(3, 0.10), (3, 0.30), (5, 0.05), (5, 0.15), (5, 0.30), (7, 0.10), (7, 0.20), (9, 0.15), (11, 0.12), (11, 0.20),
         (15, 0.15), (4, 0.2), (6, 0.15)]
out = {}
n = 0
rng = np.random.RandomState(20260501)
for L, p in cases:
    for rep in range(6 if L <= 11 else 2):
        np.random.seed(rng.randint(1 << 30))
        code = Planar(L)
        if rep == 5:                      # a layer without defects: only Z errors / only X errors / nothing at all
            kind = n % 3
            code.generate_random_error(p if kind == 1 else 0.0, 0.0, p if kind == 0 else 0.0)
        else:
            code.generate_random_error(p / 3, p / 3, p / 3)
        code.syndrom()
        m = mwpm.MWPM(code)
        w_con = np.zeros((2, 2), dtype=np.int64)
        w_free = np.zeros(2, dtype=np.int64)
        for layer in range(2):
            if np.any(m.get_layer(layer)):
                for parity in range(2):
                    w_con[layer, parity] = np.count_nonzero(m.solve_layer(layer, parity))
                w_free[layer] = np.count_nonzero(m.solve_layer(layer))
        classes = mwpm.class_sorted_mwpm(code)
        chains = np.stack([c.qubit_matrix for c in classes]).astype(np.uint8)
        for i, c in enumerate(classes):
            assert c.define_equivalence_class() == i
        k = "c%03d_" % n
        out[k + "L"] = np.int64(L)
        out[k + "qm"] = code.qubit_matrix.astype(np.uint8)
        out[k + "vertex"] = np.asarray(code.vertex_defects, dtype=np.uint8)
        out[k + "plaquette"] = np.asarray(code.plaquette_defects, dtype=np.uint8)
        out[k + "w_constrained"] = w_con
        out[k + "w_free"] = w_free
        out[k + "chains"] = chains
        out[k + "chain_weights"] = np.array([int(np.count_nonzero(c)) for c in chains], dtype=np.int64)
        out[k + "regular_class"] = np.int64(mwpm.regular_mwpm(code))
        n += 1
out["n_cases"] = np.int64(n)
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "mwpm_planar.npz")
np.savez_compressed(path, **out)
print("wrote", path, n, "cases")
