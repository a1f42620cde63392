"""CPU restatement of the reference's planar MWPM graphs -- TEST INFRASTRUCTURE ONLY (the product never imports this).

Follows /root/reference/src/mwpm.py: generate_edges (:66-133), generate_edges_constrained (:136-229) and
Planar_code.syndrom (planar_model.py:134-153).  The reference solves the graphs with the external blossom5 binary
(:376-405); here networkx.min_weight_matching does, so only solver-independent quantities come out: the weight of the
minimum perfect matching of each graph.  Pinned by tests/test_mwpm.py against tests/golden/mwpm_planar.npz, which was
produced by the unmodified reference's own graph builders (tests/golden/make_golden_mwpm.py).
"""
import numpy as np


def planar_defects(qm):
    """planar_model.py:134-153 -> (vertex_defects [L-1][L], plaquette_defects [L][L-1])"""
    yz = (qm == 2) | (qm == 3)
    vertex = (yz[0, 1:, :] ^ yz[0, :-1, :]) ^ (yz[1, :-1, :] ^ np.roll(yz[1, :-1, :], 1, axis=1))
    xy = (qm == 1) | (qm == 2)
    plaquette = (xy[0, :, 1:] ^ xy[0, :, :-1]) ^ (xy[1, :, :-1] ^ np.roll(xy[1, :, :-1], 1, axis=0))
    return vertex, plaquette


def _pairs(coords):
    e = []
    for i in range(len(coords)):
        for j in range(i + 1, len(coords)):
            e.append((i, j, int(abs(coords[i] - coords[j]).sum())))     # manhattan_path, mwpm.py:442-444
    return e


def edges_free(defects, layer, L):
    """mwpm.py:66-133 (planar branch) -> (edges, nbr_nodes)"""
    coords = np.array(np.nonzero(defects)).T
    n = len(coords)
    e = _pairs(coords)
    e += [(i + n, j + n, 0) for i in range(n) for j in range(i + 1, n)]
    for s in range(n):
        d = int(coords[s, layer]) + 1
        if not d * 2 < L:
            d = L - d
        e.append((s, s + n, d))
    return e, 2 * n


def edges_constrained(defects, layer, L, parity):
    """mwpm.py:136-229 -> (edges, nbr_nodes)"""
    coords = np.array(np.nonzero(defects)).T
    n = len(coords)
    e = _pairs(coords)
    b0 = coords[:, layer] + 1
    nearest = (b0 * 2 > L).astype(int)
    bdist = np.where(nearest == 1, L - b0, b0)
    n_anc = np.bincount(nearest, minlength=2)
    nodes = 2 * n
    if parity == 1:
        for b in range(2):
            if n_anc[b] == 0:
                e += [(s, n + (n + 1) * b, int(L - bdist[s])) for s in range(n)]
            n_anc[b] += 1
        nodes += 2
    for b in range(2):
        off = n + b * n_anc[0]
        e += [(i + off, j + off, 0) for i in range(n_anc[b]) for j in range(i + 1, n_anc[b])]
    counts = [0, 0]
    for s, b in enumerate(nearest):
        e.append((s, n + b * n_anc[0] + counts[b], int(bdist[s])))
        counts[b] += 1
    return e, nodes


def min_perfect_matching_weight(edges, nodes):
    import networkx as nx
    if nodes == 0:
        return 0
    g = nx.Graph()
    g.add_nodes_from(range(nodes))
    for a, b, w in edges:
        g.add_edge(a, b, weight=w)
    m = nx.min_weight_matching(g)
    assert 2 * len(m) == nodes, "no perfect matching"
    return int(sum(g[a][b]["weight"] for a, b in m))


def matching_weights(qm):
    """-> (w_free [2], w_constrained [2][2]) of an error chain's syndrome; 0 where a layer has no defects"""
    L = qm.shape[1]
    layers = planar_defects(qm)
    w_free, w_con = np.zeros(2, np.int64), np.zeros((2, 2), np.int64)
    for layer, d in enumerate(layers):
        if d.any():
            w_free[layer] = min_perfect_matching_weight(*edges_free(d, layer, L))
            for parity in range(2):
                w_con[layer, parity] = min_perfect_matching_weight(*edges_constrained(d, layer, L, parity))
    return w_free, w_con
