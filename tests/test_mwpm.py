"""Class-sorted MWPM start states (reference src/mwpm.py) -- host code, runs without a GPU.

A minimum-weight matching is not unique, so parity is stated on what every exact solver of the reference's graphs must
agree on: the matching weight per (layer, parity), the syndrome of every returned chain, and its class.  Golden vectors
come from the unmodified reference's own graph builders with networkx in place of the absent blossom5 binary
(tests/golden/make_golden_mwpm.py); oracle/mwpm_oracle.py restates the graphs for fresh random cases."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import mwpm_oracle as MO  # noqa: E402
from mcmc_qec_toric_rl_b200 import _lib  # noqa: E402
from mcmc_qec_toric_rl_b200.src import mwpm  # noqa: E402
from mcmc_qec_toric_rl_b200.src.planar_model import Planar_code  # noqa: E402

GOLD = np.load(os.path.join(ROOT, "tests", "golden", "mwpm_planar.npz"))
N_GOLD = int(GOLD["n_cases"])


def _z_part(c):
    return int(np.isin(c, (2, 3)).sum())


def _x_part(c):
    return int(np.isin(c, (1, 2)).sum())


def _random_error(rng, L, p):
    q = ((rng.random((2, L, L)) < p) * rng.integers(1, 4, (2, L, L))).astype(np.uint8)
    q[1, -1, :] = 0
    q[1, :, -1] = 0
    return q


def _check_chains(qm, chains):
    v, pl = MO.planar_defects(qm)
    for cls, ch in enumerate(chains):
        cv, cp = MO.planar_defects(ch)
        assert np.array_equal(cv, v) and np.array_equal(cp, pl), "a start chain does not reproduce the syndrome"
        assert int(_lib.host_planar_class(ch[None])[0]) == cls
        assert not ch[1, -1, :].any() and not ch[1, :, -1].any()


@pytest.mark.parametrize("i", range(N_GOLD))
def test_oracle_graphs_reproduce_the_reference_golden_weights(i):
    k = "c%03d_" % i
    w_free, w_con = MO.matching_weights(GOLD[k + "qm"])
    assert np.array_equal(w_free, GOLD[k + "w_free"])
    assert np.array_equal(w_con, GOLD[k + "w_constrained"])
    v, pl = MO.planar_defects(GOLD[k + "qm"])
    assert np.array_equal(v, GOLD[k + "vertex"] != 0) and np.array_equal(pl, GOLD[k + "plaquette"] != 0)


@pytest.mark.parametrize("i", range(N_GOLD))
def test_native_matching_reproduces_the_reference_golden(i):
    k = "c%03d_" % i
    L, qm = int(GOLD[k + "L"]), GOLD[k + "qm"]
    chains, w = _lib.mwpm_planar(L, qm=qm[None])
    chains2, w2 = _lib.mwpm_planar(L, vertex_defects=GOLD[k + "vertex"][None], plaquette_defects=GOLD[k + "plaquette"][None])
    assert np.array_equal(w, w2)
    has = [bool(GOLD[k + "vertex"].any()), bool(GOLD[k + "plaquette"].any())]
    for layer in range(2):
        if has[layer]:
            assert np.array_equal(w[0, layer], GOLD[k + "w_constrained"][layer])
    _check_chains(qm, chains[0])
    # per class, the Z part and the X part of the chain weigh what the reference's chains weigh (their overlap -- the
    # number of Y -- depends on which minimum matching a solver returns)
    for cls in range(4):
        assert _z_part(chains[0, cls]) == _z_part(GOLD[k + "chains"][cls])
        assert _x_part(chains[0, cls]) == _x_part(GOLD[k + "chains"][cls])
    sol, wf = _lib.mwpm_planar(L, qm=qm[None], class_sorted=False)
    assert np.array_equal(wf[0], GOLD[k + "w_free"])
    sv, sp = MO.planar_defects(sol[0])
    assert np.array_equal(sv, GOLD[k + "vertex"] != 0) and np.array_equal(sp, GOLD[k + "plaquette"] != 0)
    assert _z_part(sol[0]) == wf[0, 0] and _x_part(sol[0]) == wf[0, 1]


@pytest.mark.parametrize("L,p,n", [(5, 0.2, 40), (7, 0.15, 30), (8, 0.12, 20), (11, 0.18, 12), (15, 0.2, 6), (21, 0.2, 3)])
def test_native_matching_against_the_oracle_on_random_syndromes(L, p, n):
    rng = np.random.default_rng(1000 + L)
    qm = np.stack([_random_error(rng, L, p) for _ in range(n)])
    chains, w = _lib.mwpm_planar(L, qm=qm, threads=4)
    sols, wf = _lib.mwpm_planar(L, qm=qm, class_sorted=False, threads=1)
    for s in range(n):
        o_free, o_con = MO.matching_weights(qm[s])
        v, pl = MO.planar_defects(qm[s])
        for layer, d in enumerate((v, pl)):
            if d.any():
                assert np.array_equal(w[s, layer], o_con[layer])
        assert np.array_equal(wf[s], o_free)
        _check_chains(qm[s], chains[s])
        # the free matching is the better of the two parities
        for layer, d in enumerate((v, pl)):
            if d.any():
                assert wf[s, layer] <= w[s, layer].min()


@pytest.mark.parametrize("L,p,n", [(4, 0.25, 60), (5, 0.3, 60), (7, 0.2, 40), (9, 0.1, 30), (13, 0.2, 10), (21, 0.15, 3)])
def test_folded_graph_has_the_weights_of_the_reference_graph(L, p, n):
    """The library matches on the defects alone (border-to-border pairs as edges of weight b_i + b_j, one extra node per
    border whose count of border matches must be odd); `unreduced` solves the reference's graphs as written, one ancilla per
    defect.  Same minimum weights, valid chains in the right classes either way."""
    rng = np.random.default_rng(2000 + L)
    qm = np.stack([_random_error(rng, L, p) for _ in range(n)])
    for class_sorted in (True, False):
        a, wa = _lib.mwpm_planar(L, qm=qm, class_sorted=class_sorted)
        b, wb = _lib.mwpm_planar(L, qm=qm, class_sorted=class_sorted, unreduced=True)
        assert np.array_equal(wa, wb)
        for s in range(n):
            if class_sorted:
                _check_chains(qm[s], b[s])
                for cls in range(4):
                    assert _z_part(a[s, cls]) == _z_part(b[s, cls]) and _x_part(a[s, cls]) == _x_part(b[s, cls])
            else:
                assert _z_part(a[s]) == _z_part(b[s]) == wa[s, 0] and _x_part(a[s]) == _x_part(b[s]) == wa[s, 1]


def test_layers_without_defects_take_the_reference_logicals():
    L = 5
    chains, w = _lib.mwpm_planar(L, qm=np.zeros((1, 2, L, L), np.uint8))
    # generate_classes, mwpm.py:426-430: no defects -> [identity, apply_logical((not layer) * 2 + 1)]
    assert not chains[0, 0].any()
    assert np.array_equal(w[0], [[0, 2 * L - 1], [0, L]])
    x_row = np.zeros((2, L, L), np.uint8)
    x_row[0, 0, :] = 1
    assert np.array_equal(chains[0, 1], x_row)
    _check_chains(np.zeros((2, L, L), np.uint8), chains[0])


def test_python_mirror_of_src_mwpm():
    np.random.seed(5)
    code = Planar_code(7)
    code.generate_random_error(0.05, 0.05, 0.05)
    code.syndrom()
    classes = mwpm.class_sorted_mwpm(code)
    assert len(classes) == 4 and all(type(c) is Planar_code for c in classes)
    for i, c in enumerate(classes):
        assert c.define_equivalence_class() == i
        assert c.qubit_matrix.dtype == np.uint8 and c.qubit_matrix.shape == (2, 7, 7)
        c.syndrom()
        assert np.array_equal(c.vertex_defects, code.vertex_defects) and np.array_equal(c.plaquette_defects, code.plaquette_defects)
    reg = mwpm.regular_mwpm(code)
    w = [c.count_errors() for c in classes]
    sol = mwpm.MWPM(code).solve()
    assert 0 <= reg < 4 and np.count_nonzero(sol) <= min(w) + min(_z_part(sol), _x_part(sol))
    assert mwpm.enhanced_mwpm(code) in np.where(np.array(w) == min(w))[0]
    assert mwpm.enhanced_mwpm(code, "uncorrelated") in range(4)
    assert mwpm.enhanced_mwpm(code, "biased", p_xyz=np.array([0.01, 0.01, 0.1])) in range(4)
    with pytest.raises(AssertionError):
        from mcmc_qec_toric_rl_b200.src.toric_model import Toric_code
        mwpm.class_sorted_mwpm(Toric_code(5))
    # batched form = one call per code
    batch = mwpm.class_sorted_mwpm_batch(code.qubit_matrix[None], 7)
    assert np.array_equal(batch[0], np.stack([c.qubit_matrix for c in classes]))


def test_matching_decoders_as_a_workload():
    from mcmc_qec_toric_rl_b200 import generate_data as G
    rng = np.random.default_rng(3)
    L, S = 7, 200
    qm = np.stack([_random_error(rng, L, 0.06) for _ in range(S)])
    true = _lib.host_planar_class(qm)
    for method in ("MWPM", "eMWPM"):
        res = G.mwpm_batch(dict(code="planar", size=L, method=method), qm.reshape(S, -1), true)
        assert res["choice"].shape == (S,) and res["failures"] == int((res["choice"] != true).sum())
        assert res["failures"] < 0.2 * S       # d = 7 at p = 0.06 is far below threshold


@pytest.mark.gpu
def test_stdc_from_mwpm_start_states():
    """decoders.py:272-279 with a list of per-class codes (`mwpm_init`, generate_data.py:126-128): chains started from
    the class-constrained matchings find the true class as often as chains started from the hidden error."""
    from mcmc_qec_toric_rl_b200 import generate_data as G
    params = dict(code="planar", method="STDC", size=7, noise="depolarizing", p_error=0.08, p_sampling=0.25, droplets=8,
                  steps=7 ** 4, mwpm_init=True)
    a = G.generate_batch(params, 300, seed=11)
    b = G.generate_batch(dict(params, mwpm_init=False), 300, seed=11)
    assert np.array_equal(a["qubit"], b["qubit"]) and np.array_equal(a["eq_true"], b["eq_true"])
    assert a["distr"].shape == (300, 4) and np.allclose(a["distr"].sum(axis=1), 100.0)
    fa, fb = a["failures"], b["failures"]
    assert abs(fa - fb) <= 3 * np.sqrt(fa + fb + 1), (fa, fb)
    assert (a["choice"] == b["choice"]).mean() > 0.93
