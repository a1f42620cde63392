"""The C oracle replays every golden vector recorded from the seeded reference
(tests/golden/make_golden.py).  Integer outputs must match bit-exactly; float
outputs to 1e-9 relative (sums are taken in a different order)."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))


def load(variant):
    z = np.load(os.path.join(HERE, "golden", f"golden_{variant}.npz"))
    manifest = json.loads(str(z["manifest"]))
    cases = []
    for i, meta in enumerate(manifest):
        c = dict(meta)
        for k in z.files:
            if k.startswith(f"{i}."):
                c[k.split(".", 1)[1]] = z[k]
        cases.append(c)
    return cases


CASES = [(v, i, c) for v in ("shipped", "toric") for i, c in enumerate(load(v))]
IDS = [f"{v}-{i}-{c['kind']}-{c.get('geom')}{c.get('L')}" for v, i, c in CASES]


def streams_for(n, nb):
    return [nb] * n


@pytest.mark.parametrize("variant,idx,c", CASES, ids=IDS)
def test_golden(variant, idx, c):
    kind = c["kind"]
    g = O.GEOM[c["geom"]]
    L = c["L"]
    q = c["q"]
    if kind == "apply_stabilizer":
        for (r, col, op), want, d in zip(c["stabs"], c["out"], c["dE"]):
            got, gd = O.apply_stabilizer(g, L, q, int(r), int(col), int(op))
            assert np.array_equal(got, want) and gd == d
    elif kind == "apply_logical":
        for (op, layer, xp, zp), want, d in zip(c["args"], c["out"], c["dE"]):
            got, gd = O.apply_logical(g, L, q, int(op), int(layer), int(xp), int(zp))
            assert np.array_equal(got, want) and gd == d
    elif kind == "class":
        assert [O.eq_class(g, L, x) for x in q] == list(c["out"])
    elif kind == "to_class":
        for e in range(16):
            assert np.array_equal(O.to_class(g, L, q, e), c["out"][e])
    elif kind == "rain":
        assert np.array_equal(O.rain(g, L, q, O.Stream.mt(c["np_seed"])), c["out"])
    elif kind in ("random_stabilizer", "random_logical"):
        nb = O.Stream.mt(c["nb_seed"])
        cur = q.copy()
        for d in c["dE"]:
            if kind == "random_logical":
                cur, gd = O.apply_random_logical(g, L, cur, nb)
            else:
                flat = np.ascontiguousarray(cur).reshape(-1).copy()
                import ctypes
                r, cc, op = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
                O.lib().qo_draw_stabilizer(g, L, ctypes.c_void_p(nb.h), ctypes.byref(r), ctypes.byref(cc), ctypes.byref(op))
                cur, gd = O.apply_stabilizer(g, L, cur, r.value, cc.value, op.value)
            assert gd == d
        assert np.array_equal(cur, c["out"])
    elif kind == "update_chain":
        nb, py = O.Stream.mt(c["nb_seed"]), O.Stream.py(c["py_seed"])
        cur = q.copy()
        for want in c["out"]:
            cur = O.update_chain(g, L, cur, c["p"], c["p_logical"], c["iters"], nb, py)
            assert np.array_equal(cur, want)
    elif kind == "update_chain_weighted":
        nb, py = O.Stream.mt(c["nb_seed"]), O.Stream.py(c["py_seed"])
        cur = q.copy()
        wk = 0 if c["wkind"] == "alpha" else 1
        nx, ny, nz = [(q == k).sum() for k in (1, 2, 3)]
        ne = float(nz + c["b"] * (nx + ny))
        for want, wne in zip(c["out"], c["n_eff"]):
            cur, ne = O.update_chain_weighted(wk, g, L, cur, c["a"], c["b"], c["p_logical"], c["iters"], nb, py, ne)
            assert np.array_equal(cur, want)
            if wk == 0:
                assert ne == wne
    elif kind == "ladder":
        lk = {"dep": 0, "alpha": 1, "biased": 2}[c["lkind"]]
        nb, py = O.Stream.mt(c["nb_seed"]), O.Stream.py(c["py_seed"])
        lad = O.Ladder(lk, g, L, q, c["bottom"], c["Nc"], c["p_logical"], c["b"])
        for st, fl, t0 in zip(c["states"], c["flags"], c["tops0"]):
            lad.step(c["iters"], nb, py)
            assert np.array_equal(lad.qm.reshape(st.shape), st)
            assert list(lad.flags) == list(fl)
            assert lad.tops0.value == t0
    elif kind == "pteq":
        lk = {"dep": 0, "alpha": 1, "biased": 2}[c["lkind"]]
        nb, py = O.Stream.mt(c["nb_seed"]), O.Stream.py(c["py_seed"])
        pct, info = O.pteq(lk, g, L, q, c["p"], nb, py, param_b=c["b"], steps=c["steps"])
        assert np.array_equal(pct, c["out"]), info
    elif kind == "stdc_alpha":
        nb, py = O.Stream.mt(c["nb_seed"]), O.Stream.py(c["py_seed"])
        inits = [O.apply_logical(g, L, q, O.eq_class(g, L, q) ^ e)[0] for e in range(4)]
        out, _ = O.stdc_alpha(g, L, inits, c["pz_tilde_sampling"], c["alpha"], c["pz_tilde"], c["steps"], nb, py)
        np.testing.assert_allclose(out, c["out"], rtol=1e-9)
    elif kind == "chain_fast":
        nb = O.Stream.mt(c["nb_seed"])
        cg = O.GEOM[c["chain_geom"]]
        factor = (c["p"] / 3.0) / (1.0 - c["p"])
        cur = q.copy()
        for want in c["out"]:
            cur = O.update_chain_fast(cg, L, cur, factor, c["iters"], nb)
            assert np.array_equal(cur, want)
    elif kind in ("stdc", "strc"):
        nb, np_ = O.Stream.mt(c["nb_seed"]), O.Stream.mt(c["np_seed"])
        cg = O.GEOM[c["chain_geom"]]
        n_eq = O.neq(g)
        fn = O.stdc if kind == "stdc" else O.strc
        out = fn(g, cg, L, c["inits"], c["p_error"], c["p_sampling"], 1, c["steps"], [nb] * n_eq, [np_] * n_eq,
                 randomize=bool(c["randomize"]), conv_mult=float(c["conv_mult"]))
        np.testing.assert_allclose(out, c["out"], rtol=1e-9)
    elif kind == "single_temp":
        nb = O.Stream.mt(c["nb_seed"])
        cg = O.GEOM[c["chain_geom"]]
        out = O.single_temp(g, cg, L, c["inits"], c["p"], c["max_iters"], [nb] * O.neq(g))
        np.testing.assert_allclose(out, c["out"], rtol=1e-12)
    elif kind == "chain_fast_xyz":
        nb = O.Stream.mt(c["nb_seed"])
        cg = O.GEOM[c["chain_geom"]]
        ps = c["p_sampling"]
        factors = ps / (1.0 - ps.sum())
        cur = q.copy()
        for want in c["out"]:
            cur = O.update_chain_fast_xyz(cg, L, cur, factors, c["iters"], nb)
            assert np.array_equal(cur, want)
    elif kind == "stdc_general_noise":
        nb = O.Stream.mt(c["nb_seed"])
        cg = O.GEOM[c["chain_geom"]]
        ps = c["p_sampling"] if c["p_sampling"].size else float(c["p_xyz"].sum())
        out, out_s, _ = O.stdc_general_noise(g, cg, L, c["inits"], c["p_xyz"], ps, 1, c["steps"], [nb] * O.neq(g))
        np.testing.assert_allclose(out, c["out"], rtol=1e-9)
        np.testing.assert_allclose(out_s, c["out_shortest"], rtol=1e-9)
    elif kind in ("ptdc", "ptrc"):
        nb, py = O.Stream.mt(c["nb_seed"]), O.Stream.py(c["py_seed"])
        n_eq = O.neq(g)
        out = O.ptxc(0 if kind == "ptdc" else 1, g, L, c["inits"], c["p_error"], c["p_sampling"], 1, c["Nc"],
                     c["steps"] // c["Nc"], [nb] * n_eq, [py] * n_eq)
        # the reference truncates to uint8: allow the float result to sit within rounding of a truncation boundary
        assert np.array_equal(np.floor(out + 1e-9).astype(np.uint8), c["out"]) or \
            np.array_equal(out.astype(np.uint8), c["out"]), (out, c["out"])
    elif kind == "pteq_alpha_shortest":
        nb, py = O.Stream.mt(c["nb_seed"]), O.Stream.py(c["py_seed"])
        pct, z, sn, info = O.pteq_with_shortest(1, g, L, q, c["p"], nb, py, param_b=c["b"], steps=c["steps"])
        assert np.array_equal(pct, c["out"]), info
        np.testing.assert_allclose(z, c["out_unique"], rtol=1e-9)
        np.testing.assert_allclose(sn, c["out_shortest_n"], rtol=1e-12)
    else:
        raise AssertionError("unknown golden kind " + kind)
