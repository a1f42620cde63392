"""GPU parity tests: the CUDA path (through the C ABI) against the oracle and against the
golden vectors recorded from the reference.  Integer work (trajectories, weight changes,
accept decisions, distinct-chain histograms) must match bit-exactly in replay mode; the
float64 class distributions to 1e-9 relative (sums in a different order, device exp)."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def ctx():
    from mcmc_qec_toric_rl_b200 import _lib
    return _lib.default_context(0)


def rand_lattice(rng, g, L, p=0.15):
    shape = (2, L, L) if g in (O.TORIC, O.PLANAR) else (L, L)
    q = ((rng.random(shape) < p) * rng.integers(1, 4, shape)).astype(np.uint8)
    if g == O.PLANAR:
        q[1, -1, :] = 0
        q[1, :, -1] = 0
    return q


def golden(variant):
    z = np.load(os.path.join(HERE, "golden", f"golden_{variant}.npz"))
    manifest = json.loads(str(z["manifest"]))
    out = []
    for i, meta in enumerate(manifest):
        c = dict(meta)
        for k in z.files:
            if k.startswith(f"{i}."):
                c[k.split(".", 1)[1]] = z[k]
        out.append(c)
    return out


# ------------------------------------------------------------------ replay: chains
@pytest.mark.parametrize("g,L", [(O.TORIC, 5), (O.TORIC, 15), (O.TORIC, 21), (O.PLANAR, 5), (O.PLANAR, 7),
                                 (O.PLANAR, 21), (O.ROTATED, 5), (O.ROTATED, 25), (O.XZZX, 7), (O.XZZX, 21)])
def test_replay_chain_matches_oracle(ctx, g, L):
    rng = np.random.default_rng(1000 + 10 * g + L)
    chains, iters = 48, 300
    k = 3 if g in (O.TORIC, O.PLANAR) else 5
    p = 0.3
    factor = (p / 3.0) / (1.0 - p)
    qm0 = np.stack([rand_lattice(rng, g, L).reshape(-1) for _ in range(chains)])
    u = rng.random((chains, iters, k + 1))
    fin, dE, acc, traj = ctx.replay_chain(g, L, qm0, u, p, want_traj=True)
    for ch in range(chains):
        want, wdE, wacc = O.update_chain_fast(g, L, qm0[ch], factor, iters, O.Stream.replay(u[ch].reshape(-1)), trace=True)
        assert np.array_equal(dE[ch], wdE)
        assert np.array_equal(acc[ch], wacc)
        assert np.array_equal(fin[ch], want)
        assert np.array_equal(traj[ch, -1], want)


@pytest.mark.parametrize("variant", ["shipped", "toric"])
def test_replay_chain_matches_reference_golden(ctx, variant):
    """Feed the reference's own numba MT19937 draws: trajectories equal the recorded reference run."""
    n = 0
    for c in golden(variant):
        if c["kind"] != "chain_fast":
            continue
        g = O.GEOM[c["chain_geom"]]
        L, iters, blocks = c["L"], c["iters"], c["blocks"]
        u = np.random.RandomState(c["nb_seed"]).random_sample(iters * blocks * 4).reshape(1, iters * blocks, 4)
        qm0 = c["q"].reshape(1, -1).copy()
        fin, dE, acc, traj = ctx.replay_chain(g, L, qm0, u, c["p"], want_traj=True)
        snaps = traj[0, iters - 1::iters]
        assert np.array_equal(snaps, c["out"].reshape(blocks, -1))
        n += 1
    assert n >= 1


# ------------------------------------------------------------------ replay: STDC
def _stdc_streams(rng, n_chains, steps, iters, k, L):
    u_nb = rng.random((n_chains, steps * iters, k + 1))
    u_np = rng.random((n_chains, 2 * L * L))
    return u_nb, u_np


@pytest.mark.parametrize("gcode,gchain,L,droplets,per_class", [
    (O.TORIC, O.TORIC, 5, 1, False), (O.TORIC, O.TORIC, 5, 3, False), (O.TORIC, O.PLANAR, 5, 2, False),
    (O.TORIC, O.TORIC, 7, 2, False), (O.PLANAR, O.PLANAR, 5, 3, True), (O.PLANAR, O.PLANAR, 5, 2, False),
    (O.TORIC, O.TORIC, 17, 2, False), (O.ROTATED, O.ROTATED, 5, 2, False), (O.XZZX, O.XZZX, 5, 2, True)])
def test_stdc_replay_matches_oracle(ctx, gcode, gchain, L, droplets, per_class):
    rng = np.random.default_rng(77 + L + droplets)
    S, steps, iters = 3, 120, 5
    n_eq = O.neq(gcode)
    k = 3 if gchain in (O.TORIC, O.PLANAR) else 5
    randomize = gcode in (O.TORIC, O.PLANAR) and not per_class
    qs = [rand_lattice(rng, gcode, L, 0.1) for _ in range(S)]
    if per_class:
        qm = np.stack([O.all_classes(gcode, L, q) for q in qs])
    else:
        qm = np.stack([q.reshape(-1) for q in qs])
    n_chains = S * n_eq * droplets
    u_nb, u_np = _stdc_streams(rng, n_chains, steps, iters, k, L)
    out, st, hist = ctx.stdc(gcode, gchain, L, qm, 0.1, 0.25, droplets, steps, iters=iters, per_class=per_class,
                             randomize=randomize, u_nb=u_nb, u_np=u_np, want_hist=True)
    for s in range(S):
        inits = qm[s] if per_class else O.all_classes(gcode, L, qs[s])
        base = s * n_eq * droplets
        nb = [O.Stream.replay(u_nb[base + i].reshape(-1)) for i in range(n_eq * droplets)]
        np_ = [O.Stream.replay(u_np[base + i]) for i in range(n_eq * droplets)]
        want, wdist, whist = O.stdc(gcode, gchain, L, inits, 0.1, 0.25, droplets, steps, nb, np_, iters=iters,
                                    randomize=randomize, want_hist=True)
        assert np.array_equal(hist[s].astype(np.int64), whist), f"N(n) histogram differs for syndrome {s}"
        np.testing.assert_allclose(out[s], want, rtol=1e-9)
    assert st["distinct"] == int(hist.sum())


def _golden_class_streams(fn, c):
    """The reference ran its classes one after the other on ONE numba stream and ONE numpy stream (droplets=1).
    With the conv_mult early stop a class consumes a data-dependent number of draws, so the start of every class's
    slice is found by replaying the classes through the oracle (pinned to these very vectors) and reading the
    stream positions."""
    gcode, gchain, L = O.GEOM[c["geom"]], O.GEOM[c["chain_geom"]], c["L"]
    n_eq, steps, iters = O.neq(gcode), c["steps"], 5
    inits = c["inits"].reshape(n_eq, -1)
    total = n_eq * steps * iters * 4
    full_nb = np.random.RandomState(c["nb_seed"]).random_sample(total + steps * iters * 4)
    full_np = np.random.RandomState(c["np_seed"]).random_sample((n_eq + 1) * 2 * L * L)
    nb, np_ = O.Stream.replay(full_nb), O.Stream.replay(full_np)
    u_nb = np.zeros((n_eq, steps * iters, 4))
    u_np = np.zeros((n_eq, 2 * L * L))
    for e in range(n_eq):
        a, b = nb.drawn, np_.drawn
        u_nb[e] = full_nb[a:a + steps * iters * 4].reshape(-1, 4)
        u_np[e] = full_np[b:b + 2 * L * L]
        getattr(O, fn)(gcode, gchain, L, inits[e:e + 1], c["p_error"], c["p_sampling"], 1, steps, [nb], [np_],
                       randomize=bool(c["randomize"]), conv_mult=float(c["conv_mult"]))
    return u_nb, u_np


def _golden_driver(ctx, fn, variant):
    n = 0
    for c in golden(variant):
        if c["kind"] != fn:
            continue
        gcode, gchain, L = O.GEOM[c["geom"]], O.GEOM[c["chain_geom"]], c["L"]
        n_eq, steps = O.neq(gcode), c["steps"]
        per_class = gcode != O.TORIC
        u_nb, u_np = _golden_class_streams(fn, c)
        qm = (c["inits"].reshape(1, n_eq, -1) if per_class else c["q"].reshape(1, -1)).copy()
        out, st = getattr(ctx, fn)(gcode, gchain, L, qm, c["p_error"], c["p_sampling"], 1, steps, iters=5, per_class=per_class,
                                   randomize=bool(c["randomize"]), conv_mult=float(c["conv_mult"]), u_nb=u_nb, u_np=u_np)
        np.testing.assert_allclose(out[0], c["out"], rtol=1e-9)
        n += 1
    return n


@pytest.mark.parametrize("variant", ["shipped", "toric"])
def test_stdc_replay_matches_reference_golden(ctx, variant):
    """STDC(droplets=1) of the seeded reference, with and without conv_mult: same class distribution from its own draws."""
    assert _golden_driver(ctx, "stdc", variant) >= 4


# ------------------------------------------------------------------ replay: STRC, single_temp
@pytest.mark.parametrize("gcode,gchain,L,droplets,per_class", [
    (O.TORIC, O.TORIC, 5, 1, False), (O.TORIC, O.TORIC, 5, 4, False), (O.TORIC, O.PLANAR, 5, 2, False),
    (O.PLANAR, O.PLANAR, 5, 3, True), (O.PLANAR, O.PLANAR, 7, 2, False), (O.TORIC, O.TORIC, 17, 2, False),
    (O.ROTATED, O.ROTATED, 5, 2, True)])
def test_strc_replay_matches_oracle(ctx, gcode, gchain, L, droplets, per_class):
    rng = np.random.default_rng(177 + L + droplets)
    S, steps, iters = 3, 150, 5
    n_eq = O.neq(gcode)
    k = 3 if gchain in (O.TORIC, O.PLANAR) else 5
    randomize = gcode in (O.TORIC, O.PLANAR) and not per_class
    qs = [rand_lattice(rng, gcode, L, 0.1) for _ in range(S)]
    qm = np.stack([O.all_classes(gcode, L, q) for q in qs]) if per_class else np.stack([q.reshape(-1) for q in qs])
    n_chains = S * n_eq * droplets
    u_nb, u_np = _stdc_streams(rng, n_chains, steps, iters, k, L)
    out, st, mh, info = ctx.strc(gcode, gchain, L, qm, 0.1, 0.25, droplets, steps, iters=iters, per_class=per_class,
                                 randomize=randomize, u_nb=u_nb, u_np=u_np, want_hist=True)
    for s in range(S):
        inits = qm[s] if per_class else O.all_classes(gcode, L, qs[s])
        base = s * n_eq * droplets
        nb = [O.Stream.replay(u_nb[base + i].reshape(-1)) for i in range(n_eq * droplets)]
        np_ = [O.Stream.replay(u_np[base + i]) for i in range(n_eq * droplets)]
        want, wmh, winfo = O.strc(gcode, gchain, L, inits, 0.1, 0.25, droplets, steps, nb, np_, iters=iters,
                                  randomize=randomize, want_hist=True)
        assert np.array_equal(mh[s].astype(np.int64), wmh), f"m(n) histogram differs for syndrome {s}"
        assert np.array_equal(info[s].astype(np.int64), winfo), f"shortest / next-shortest bookkeeping differs ({s})"
        np.testing.assert_allclose(out[s], want, rtol=1e-9)


@pytest.mark.parametrize("variant", ["shipped", "toric"])
def test_strc_replay_matches_reference_golden(ctx, variant):
    assert _golden_driver(ctx, "strc", variant) >= 4


@pytest.mark.parametrize("variant", ["shipped", "toric"])
def test_single_temp_replay_matches_reference_golden(ctx, variant):
    n = 0
    for c in golden(variant):
        if c["kind"] != "single_temp":
            continue
        gcode, gchain, L = O.GEOM[c["geom"]], O.GEOM[c["chain_geom"]], c["L"]
        n_eq, m = O.neq(gcode), c["max_iters"]
        u_nb = np.random.RandomState(c["nb_seed"]).random_sample(n_eq * m * 5 * 4).reshape(n_eq, m * 5, 4)
        out, _ = ctx.single_temp(gcode, gchain, L, c["inits"].reshape(1, n_eq, -1).copy(), c["p"], m, per_class=True, u_nb=u_nb)
        np.testing.assert_allclose(out[0], c["out"], rtol=1e-12)
        n += 1
    assert n >= 2


@pytest.mark.parametrize("g,L", [(O.TORIC, 5), (O.PLANAR, 7), (O.TORIC, 19)])
def test_single_temp_replay_matches_oracle(ctx, g, L):
    rng = np.random.default_rng(300 + L)
    S, max_iters, iters = 4, 90, 5
    n_eq = O.neq(g)
    qs = [rand_lattice(rng, g, L, 0.1) for _ in range(S)]
    qm = np.stack([q.reshape(-1) for q in qs])
    u_nb = rng.random((S * n_eq, max_iters * iters, 4))
    out, st = ctx.single_temp(g, g, L, qm, 0.2, max_iters, iters=iters, u_nb=u_nb)
    for s in range(S):
        nb = [O.Stream.replay(u_nb[s * n_eq + e].reshape(-1)) for e in range(n_eq)]
        want = O.single_temp(g, g, L, O.all_classes(g, L, qs[s]), 0.2, max_iters, nb, iters=iters)
        assert np.array_equal(out[s], want)


@pytest.mark.parametrize("droplets", [1, 3])
@pytest.mark.parametrize("fn", ["stdc", "strc"])
@pytest.mark.parametrize("gcode,L", [(O.TORIC, 5), (O.PLANAR, 7), (O.ROTATED, 5)])
def test_conv_mult_early_stop_matches_oracle(ctx, fn, gcode, L, droplets):
    """conv_mult != 0 (decoders.py:257-263, :795): "new" is new to the droplet, so every chain stops at the same sample
    as the oracle's and the union of what the droplets of a class saw gives identical histograms -- also with several
    droplets per class (a set per chain + a log of the keys new to it, then the dedupe kernel)."""
    if fn == "strc" and gcode == O.ROTATED:
        pytest.skip("covered by stdc")
    rng = np.random.default_rng(4100 + L + droplets)
    S, steps, iters, conv = 3, 400, 5, 2.0
    n_eq = O.neq(gcode)
    k = 3 if gcode in (O.TORIC, O.PLANAR) else 5
    randomize = gcode in (O.TORIC, O.PLANAR)
    qs = [rand_lattice(rng, gcode, L, 0.08) for _ in range(S)]
    qm = np.stack([q.reshape(-1) for q in qs])
    u_nb, u_np = _stdc_streams(rng, S * n_eq * droplets, steps, iters, k, L)
    run = getattr(ctx, fn)
    res = run(gcode, gcode, L, qm, 0.1, 0.25, droplets, steps, iters=iters, randomize=randomize, conv_mult=conv, u_nb=u_nb,
              u_np=u_np, want_hist=True)
    assert res[1]["metropolis_steps"] < S * n_eq * droplets * steps * iters      # some chains did stop early
    for s in range(S):
        base = s * n_eq * droplets
        nb = [O.Stream.replay(u_nb[base + i].reshape(-1)) for i in range(n_eq * droplets)]
        np_ = [O.Stream.replay(u_np[base + i]) for i in range(n_eq * droplets)]
        want = getattr(O, fn)(gcode, gcode, L, O.all_classes(gcode, L, qs[s]), 0.1, 0.25, droplets, steps, nb, np_, iters=iters,
                              randomize=randomize, conv_mult=conv, want_hist=True)
        assert np.array_equal(res[2][s].astype(np.int64), want[2 if fn == "stdc" else 1])
        np.testing.assert_allclose(res[0][s], want[0], rtol=1e-9)


# ------------------------------------------------------------------ native Philox
def test_native_chain_statistics(ctx):
    """Native Philox chains: acceptance rate and mean weight agree with the oracle's MT19937 chains
    within 5 standard errors (independent chains, same start, same number of steps)."""
    g, L, p = O.TORIC, 7, 0.25
    rng = np.random.default_rng(5)
    chains, iters = 4096, 400
    q0 = rand_lattice(rng, g, L, 0.1).reshape(-1)
    qm = np.tile(q0, (chains, 1))
    st = ctx.chain_update(g, L, qm, p, iters, seed=1234)
    gpu_w = (qm != 0).sum(1)
    factor = (p / 3.0) / (1.0 - p)
    n_or = 512
    or_w, or_acc = [], 0
    for i in range(n_or):
        f, dE, acc = O.update_chain_fast(g, L, q0, factor, iters, O.Stream.mt(1000 + i), trace=True)
        or_w.append((f != 0).sum())
        or_acc += acc.sum()
    or_w = np.array(or_w)
    se = np.sqrt(gpu_w.var() / chains + or_w.var() / n_or)
    assert abs(gpu_w.mean() - or_w.mean()) < 5 * se, (gpu_w.mean(), or_w.mean(), se)
    r_gpu = st["accepted"] / (chains * iters)
    r_or = or_acc / (n_or * iters)
    assert abs(r_gpu - r_or) < 0.01, (r_gpu, r_or)
    # syndrome is preserved: every chain stays in the start's class orbit (toric stabilizers)
    assert all(O.eq_class(g, L, x) == O.eq_class(g, L, q0) for x in qm[:64])


def test_native_stdc_agrees_with_oracle(ctx):
    """Native STDC on toric d=5: class distributions close to the oracle's (different RNG, same algorithm)
    and the same most-likely class on nearly every syndrome."""
    g, L = O.TORIC, 5
    rng = np.random.default_rng(11)
    S, droplets, steps = 24, 4, 2000
    qm = np.stack([rand_lattice(rng, g, L, 0.08).reshape(-1) for _ in range(S)])
    out, st = ctx.stdc(g, g, L, qm, 0.08, 0.25, droplets, steps, seed=99)
    want = O.stdc_batch(g, g, L, qm, 0.08, 0.25, droplets, steps, seed=7, threads=8)
    assert np.allclose(out.sum(1), 100.0)
    agree = (out.argmax(1) == want.argmax(1)).mean()
    assert agree >= 0.9, agree
    assert np.abs(out - want).max() < 12.0, np.abs(out - want).max()
    assert np.abs(out - want).mean() < 1.0
    assert st["metropolis_steps"] == S * 16 * droplets * steps * 5


def test_errors_are_loud(ctx):
    from mcmc_qec_toric_rl_b200 import _lib
    with pytest.raises(TypeError):
        ctx.chain_update(O.TORIC, 5, np.zeros((1, 50), np.int64), 0.1, 5)
    bad = np.full((1, 50), 7, np.uint8)
    with pytest.raises(_lib.QecmcError):
        ctx.chain_update(O.TORIC, 5, bad, 0.1, 5)
    with pytest.raises(_lib.QecmcError):
        ctx.chain_update(O.ROTATED, 6, np.zeros((1, 36), np.uint8), 0.1, 5)


# ------------------------------------------------------------------ replay: tempering ladders
def _ladder_budget(g, Nc, iters, steps):
    k = 3 if g in (O.TORIC, O.PLANAR) else 5
    return steps * (Nc * k * iters + 6 * iters + Nc), steps * (Nc * iters + 2 * iters + Nc)


LADDER_CASES = [
    (0, O.TORIC, 5, 5, 0.5), (0, O.PLANAR, 5, 4, 0.5), (0, O.ROTATED, 7, 7, 0.5), (0, O.XZZX, 5, 5, 0.5),
    (0, O.TORIC, 7, 1, 0.5), (0, O.ROTATED, 25, 25, 0.5), (0, O.TORIC, 5, 3, 0.0),
    (1, O.TORIC, 5, 5, 0.5), (1, O.XZZX, 7, 6, 0.5), (1, O.PLANAR, 5, 1, 0.0), (1, O.ROTATED, 5, 5, 0.3),
    (2, O.XZZX, 7, 7, 0.5), (2, O.TORIC, 5, 4, 0.5), (2, O.PLANAR, 7, 2, 0.0), (2, O.XZZX, 21, 21, 0.5),
]


@pytest.mark.parametrize("kind,g,L,Nc,p_logical", LADDER_CASES)
def test_ladder_replay_matches_oracle(ctx, kind, g, L, Nc, p_logical):
    """Ladder.step on the device == the oracle's, from the same NB / PY uniform streams: rung states, flags,
    tops0 after every step and the rung-owned n_eff at the end, all bit-exact."""
    rng = np.random.default_rng(9000 + 100 * kind + 10 * g + L + Nc)
    S, steps, iters = 3, 12, 10
    bottom = 0.12 if kind != 1 else 0.2
    b = {0: 0.0, 1: 1.7, 2: 8.0}[kind]
    qs = [rand_lattice(rng, g, L, 0.12) for _ in range(S)]
    qm = np.stack([q.reshape(-1) for q in qs])
    n_nb, n_py = _ladder_budget(g, Nc, iters, steps)
    u_nb, u_py = rng.random((S, n_nb)), rng.random((S, n_py))
    out = ctx.ladder_run(g, L, kind, qm, bottom, Nc, steps, iters=iters, param_b=b, p_logical=p_logical, u_nb=u_nb, u_py=u_py,
                         snapshots=True)
    for s in range(S):
        lad = O.Ladder(kind, g, L, qs[s], bottom, Nc, p_logical, b)
        nb, py = O.Stream.replay(u_nb[s]), O.Stream.replay(u_py[s])
        for t in range(steps):
            lad.step(iters, nb, py)
            assert np.array_equal(out["snap_states"][s, t], lad.qm), f"rung states differ (ladder {s}, step {t})"
            assert np.array_equal(out["snap_flags"][s, t], lad.flags), f"flags differ (ladder {s}, step {t})"
            assert out["snap_tops0"][s, t] == lad.tops0.value
        assert np.array_equal(out["rung_states"][s], lad.qm)
        assert out["tops0"][s] == lad.tops0.value
        if kind == 1:
            assert np.array_equal(out["n_eff"][s], lad.n_eff)


def _py_uniforms(seed, n):
    import random
    r = random.Random(seed)
    return np.array([r.random() for _ in range(n)])


def test_ladder_replay_matches_reference_golden(ctx):
    """Ladder / Ladder_alpha / Ladder_biased of the seeded reference, from its own numba and CPython streams."""
    n = 0
    for c in golden("shipped"):
        if c["kind"] != "ladder":
            continue
        kind = {"dep": 0, "alpha": 1, "biased": 2}[c["lkind"]]
        g, L, Nc, iters = O.GEOM[c["geom"]], c["L"], c["Nc"], c["iters"]
        steps = len(c["tops0"])
        n_nb, n_py = _ladder_budget(g, Nc, iters, steps)
        u_nb = np.random.RandomState(c["nb_seed"]).random_sample(n_nb).reshape(1, -1)
        u_py = _py_uniforms(c["py_seed"], n_py).reshape(1, -1)
        out = ctx.ladder_run(g, L, kind, c["q"].reshape(1, -1).copy(), c["bottom"], Nc, steps, iters=iters, param_b=c["b"],
                             p_logical=c["p_logical"], u_nb=u_nb, u_py=u_py, snapshots=True)
        for t in range(steps):
            assert np.array_equal(out["snap_states"][0, t], c["states"][t].reshape(Nc, -1)), (c["geom"], c["lkind"], t)
            assert list(out["snap_flags"][0, t]) == list(c["flags"][t])
            assert out["snap_tops0"][0, t] == c["tops0"][t]
        n += 1
    assert n >= 12


PTEQ_CASES = [(0, O.TORIC, 5, 0.1, 0.0), (0, O.ROTATED, 7, 0.12, 0.0), (0, O.PLANAR, 5, 0.1, 0.0), (0, O.XZZX, 5, 0.1, 0.0),
              (1, O.XZZX, 5, 0.15, 2.0), (1, O.TORIC, 5, 0.15, 1.5), (2, O.XZZX, 7, 0.1, 10.0), (2, O.ROTATED, 5, 0.1, 3.0)]


@pytest.mark.parametrize("kind,g,L,bottom,b", PTEQ_CASES)
def test_pteq_replay_matches_oracle(ctx, kind, g, L, bottom, b):
    """PTEQ with its convergence criterion: same class percentages, same number of steps, same tops0."""
    rng = np.random.default_rng(500 + 10 * kind + g + L)
    S, iters, cap = 3, 10, 3000
    qs = [rand_lattice(rng, g, L, 0.1) for _ in range(S)]
    qm = np.stack([q.reshape(-1) for q in qs])
    n_nb, n_py = _ladder_budget(g, L, iters, cap)
    u_nb, u_py = rng.random((S, n_nb)), rng.random((S, n_py))
    pct, info = ctx.pteq(g, L, kind, qm, bottom, param_b=b, steps=cap, iters=iters, u_nb=u_nb, u_py=u_py)
    for s in range(S):
        want, winfo = O.pteq(kind, g, L, qs[s], bottom, O.Stream.replay(u_nb[s]), O.Stream.replay(u_py[s]), param_b=b,
                             steps=cap, iters=iters)
        assert info["steps"][s] == winfo["steps"], (info["steps"][s], winfo["steps"])
        assert info["since_burn"][s] == winfo["since_burn"] and info["tops0"][s] == winfo["tops0"]
        assert np.array_equal(info["counts"][s], winfo["counts"])
        assert np.array_equal(pct[s], want)


@pytest.mark.parametrize("kind,g,L,bottom,b,Nc", [(0, O.ROTATED, 9, 0.15, 0.0, None), (0, O.TORIC, 5, 0.1, 0.0, None),
                                                  (0, O.PLANAR, 7, 0.12, 0.0, 3), (0, O.ROTATED, 5, 0.1, 0.0, 2),
                                                  (2, O.XZZX, 9, 0.15, 30.0, None), (2, O.XZZX, 7, 0.7, 4.0, None),
                                                  (1, O.XZZX, 7, 0.17, 0.65, None)])
def test_native_swap_sweep_equals_the_pair_by_pair_walk(ctx, kind, g, L, bottom, b, Nc, monkeypatch):
    """Native ladders take the swap decisions of a sweep on all lanes (largest swapping exponent per pair from the power
    table, compare-and-select walk, rungs from the swap mask).  debug_set("serial_sweep") makes the same build walk the pairs
    one by one as mcmc.py:96-103 does (the replay path's code): same seed, same draws => identical results.  The 0.7
    bottom rung gives the biased ladder tables close to 1, i.e. draws beyond the tabulated exponents."""
    rng = np.random.default_rng(900 + 7 * kind + g + L)
    S = 200
    qm = np.stack([rand_lattice(rng, g, L, 0.12).reshape(-1) for _ in range(S)])
    kw = dict(param_b=b, steps=300, conv=False, seed=5, p_logical=0.5)
    if Nc:
        kw["Nc"] = Nc
    pct, info = ctx.pteq(g, L, kind, qm, bottom, **kw)
    ctx.debug_set("serial_sweep", 1)
    try:
        pct1, info1 = ctx.pteq(g, L, kind, qm, bottom, **kw)
    finally:
        ctx.debug_set("serial_sweep", -1)
    assert np.array_equal(pct, pct1)
    for k in ("steps", "since_burn", "tops0", "counts"):
        assert np.array_equal(info[k], info1[k]), k
    assert info["stats"]["accepted"] == info1["stats"]["accepted"]
    assert info["tops0"].sum() > 0 or kind == 0   # replicas did travel the ladder


def test_pteq_replay_matches_reference_golden(ctx):
    n = 0
    for c in golden("shipped"):
        if c["kind"] != "pteq":
            continue
        kind = {"dep": 0, "alpha": 1, "biased": 2}[c["lkind"]]
        g, L = O.GEOM[c["geom"]], c["L"]
        # the oracle (pinned to this very vector in test_oracle_golden) tells how many steps the reference ran
        _, winfo = O.pteq(kind, g, L, c["q"], c["p"], O.Stream.mt(c["nb_seed"]), O.Stream.py(c["py_seed"]), param_b=c["b"],
                          steps=c["steps"])
        used = int(winfo["steps"])
        n_nb, n_py = _ladder_budget(g, L, 10, used)
        u_nb = np.random.RandomState(c["nb_seed"]).random_sample(n_nb).reshape(1, -1)
        u_py = _py_uniforms(c["py_seed"], n_py).reshape(1, -1)
        pct, info = ctx.pteq(g, L, kind, c["q"].reshape(1, -1).copy(), c["p"], param_b=c["b"], steps=used, u_nb=u_nb, u_py=u_py)
        assert info["steps"][0] == used
        assert np.array_equal(pct[0], c["out"]), (c["geom"], c["lkind"])
        n += 1
    assert n >= 12


@pytest.mark.parametrize("g,L", [(O.PLANAR, 5), (O.ROTATED, 5), (O.XZZX, 7), (O.TORIC, 5)])
def test_stdc_alpha_replay_matches_oracle(ctx, g, L):
    """EWD-style STDC_Nall_n_alpha: distinct counts equal, class distribution to 1e-9."""
    rng = np.random.default_rng(640 + g + L)
    S, steps, iters, alpha = 2, 400, 5, 1.6
    n_eq = O.neq(g)
    k = 3 if g in (O.TORIC, O.PLANAR) else 5
    qs = [rand_lattice(rng, g, L, 0.1) for _ in range(S)]
    qm = np.stack([q.reshape(-1) for q in qs])
    u_nb, u_py = rng.random((S * n_eq, steps * iters * k)), rng.random((S * n_eq, steps * iters))
    out, distinct, st = ctx.stdc_alpha(g, L, qm, 0.25, alpha, 0.1, steps, iters=iters, u_nb=u_nb, u_py=u_py)
    for s in range(S):
        inits = O.all_classes(g, L, qs[s])
        nb = O.Stream.replay(u_nb[s * n_eq:(s + 1) * n_eq].reshape(-1))
        py = O.Stream.replay(u_py[s * n_eq:(s + 1) * n_eq].reshape(-1))
        want, wdist = O.stdc_alpha(g, L, inits, 0.25, alpha, 0.1, steps, nb, py, iters=iters)
        assert np.array_equal(distinct[s], wdist)
        np.testing.assert_allclose(out[s], want, rtol=1e-9)


def test_stdc_alpha_replay_matches_reference_golden(ctx):
    n = 0
    for c in golden("shipped"):
        if c["kind"] != "stdc_alpha":
            continue
        g, L, steps = O.GEOM[c["geom"]], c["L"], c["steps"]
        k = 3 if g in (O.TORIC, O.PLANAR) else 5
        u_nb = np.random.RandomState(c["nb_seed"]).random_sample(4 * steps * 5 * k).reshape(4, -1)
        u_py = _py_uniforms(c["py_seed"], 4 * steps * 5).reshape(4, -1)
        out, distinct, st = ctx.stdc_alpha(g, L, c["q"].reshape(1, -1).copy(), c["pz_tilde_sampling"], c["alpha"], c["pz_tilde"],
                                           steps, u_nb=u_nb, u_py=u_py)
        np.testing.assert_allclose(out[0], c["out"], rtol=1e-9)
        n += 1
    assert n >= 3


def test_native_pteq_agrees_with_oracle(ctx):
    """Native Philox PTEQ (fixed number of steps): the class histogram of the bottom rung agrees with the
    oracle's MT19937 run on the same syndromes -- same most likely class, percentages within sampling noise."""
    g, L, p = O.ROTATED, 5, 0.1
    rng = np.random.default_rng(21)
    S, steps = 16, 20000
    qs = [rand_lattice(rng, g, L, 0.1) for _ in range(S)]
    qm = np.stack([q.reshape(-1) for q in qs])
    pct, info = ctx.pteq(g, L, 0, qm, p, steps=steps, conv=False, seed=5)
    agree, diffs = 0, []
    for s in range(S):
        want, _ = O.pteq(0, g, L, qs[s], p, O.Stream.mt(100 + s), O.Stream.py(200 + s), steps=steps, conv=False)
        agree += int(pct[s].argmax() == want.argmax())
        diffs.append(np.abs(pct[s].astype(int) - want.astype(int)))
    diffs = np.array(diffs)
    # two independent oracle runs of this configuration differ by up to ~15 points on near-degenerate syndromes
    # (the class histogram of a 20000-step ladder has only a few hundred independent samples)
    assert diffs.max() <= 25, diffs.max()
    assert diffs.mean() <= 3.0, diffs.mean()
    assert agree >= S - 2
    assert (info["steps"] == steps).all()


# ------------------------------------------------------------------ replay: PTDC
def _ptxc_streams(rng, g, n_ladders, Nc, iters, steps):
    n_nb, n_py = _ladder_budget(g, Nc, iters, steps)
    return rng.random((n_ladders, n_nb)), rng.random((n_ladders, n_py))


@pytest.mark.parametrize("g,L,Nc,droplets,per_class", [(O.TORIC, 5, 4, 1, False), (O.TORIC, 5, 3, 2, False),
                                                        (O.PLANAR, 5, 5, 2, True), (O.ROTATED, 5, 4, 1, False),
                                                        (O.XZZX, 7, 3, 2, True)])
def test_ptdc_replay_matches_oracle(ctx, g, L, Nc, droplets, per_class):
    rng = np.random.default_rng(7100 + g + L + Nc)
    S, steps, iters = 2, 60, 10
    n_eq = O.neq(g)
    qs = [rand_lattice(rng, g, L, 0.1) for _ in range(S)]
    qm = np.stack([O.all_classes(g, L, q) for q in qs]) if per_class else np.stack([q.reshape(-1) for q in qs])
    u_nb, u_py = _ptxc_streams(rng, g, S * n_eq * droplets, Nc, iters, steps)
    out, st = ctx.ptdc(g, L, qm, 0.1, 0.25, droplets, Nc, steps, iters=iters, per_class=per_class, u_nb=u_nb, u_py=u_py)
    for s in range(S):
        base = s * n_eq * droplets
        nb = [O.Stream.replay(u_nb[base + i]) for i in range(n_eq * droplets)]
        py = [O.Stream.replay(u_py[base + i]) for i in range(n_eq * droplets)]
        want = O.ptxc(0, g, L, O.all_classes(g, L, qs[s]), 0.1, 0.25, droplets, Nc, steps, nb, py, iters=iters)
        np.testing.assert_allclose(out[s], want, rtol=1e-9)


def _golden_ladder_class_streams(c, n_eq, Nc, steps):
    """PTDC / PTRC of the reference (droplets=1): the classes ran one after the other on ONE numba and ONE CPython
    stream, each consuming a data-dependent amount (swap draws); the per-class slices start where the oracle,
    pinned to these vectors, says the previous class stopped."""
    g = O.GEOM[c["geom"]]
    n_nb, n_py = _ladder_budget(g, Nc, 10, steps)
    full_nb = np.random.RandomState(c["nb_seed"]).random_sample(n_nb * (n_eq + 1))
    full_py = _py_uniforms(c["py_seed"], n_py * (n_eq + 1))
    nb, py = O.Stream.replay(full_nb), O.Stream.replay(full_py)
    u_nb, u_py = np.zeros((n_eq, n_nb)), np.zeros((n_eq, n_py))
    inits = c["inits"].reshape(n_eq, -1)
    for e in range(n_eq):
        a, b = nb.drawn, py.drawn
        u_nb[e], u_py[e] = full_nb[a:a + n_nb], full_py[b:b + n_py]
        O.ptxc(0, g, c["L"], inits[e:e + 1], c["p_error"], c["p_sampling"], 1, Nc, steps, [nb], [py])
    return u_nb, u_py


def test_ptdc_replay_matches_reference_golden(ctx):
    n = 0
    for c in golden("shipped"):
        if c["kind"] != "ptdc":
            continue
        g, L, Nc = O.GEOM[c["geom"]], c["L"], c["Nc"]
        n_eq, steps = O.neq(g), c["steps"] // c["Nc"]
        u_nb, u_py = _golden_ladder_class_streams(c, n_eq, Nc, steps)
        out, _ = ctx.ptdc(g, L, c["inits"].reshape(1, n_eq, -1).copy(), c["p_error"], c["p_sampling"], 1, Nc, steps, per_class=True,
                          u_nb=u_nb, u_py=u_py)
        assert np.array_equal(out[0].astype(np.uint8), c["out"]) or np.array_equal(np.floor(out[0] + 1e-9).astype(np.uint8), c["out"])
        n += 1
    assert n >= 2


# ------------------------------------------------------------------ replay: PTRC
@pytest.mark.parametrize("g,L,Nc,droplets,per_class", [(O.TORIC, 5, 4, 1, False), (O.PLANAR, 5, 3, 2, True),
                                                        (O.ROTATED, 5, 4, 2, False), (O.TORIC, 7, 2, 3, False)])
def test_ptrc_replay_matches_oracle(ctx, g, L, Nc, droplets, per_class):
    """Per-rung N(n) and m(n) (droplets summed) bit-exact, the class distribution to 1e-9."""
    rng = np.random.default_rng(7300 + g + L + Nc)
    S, steps, iters = 2, 60, 10
    n_eq = O.neq(g)
    qs = [rand_lattice(rng, g, L, 0.1) for _ in range(S)]
    qm = np.stack([O.all_classes(g, L, q) for q in qs]) if per_class else np.stack([q.reshape(-1) for q in qs])
    u_nb, u_py = _ptxc_streams(rng, g, S * n_eq * droplets, Nc, iters, steps)
    out, st, Nh, mh = ctx.ptrc(g, L, qm, 0.1, 0.25, droplets, Nc, steps, iters=iters, per_class=per_class, u_nb=u_nb, u_py=u_py,
                               want_hist=True)
    for s in range(S):
        base = s * n_eq * droplets
        nb = [O.Stream.replay(u_nb[base + i]) for i in range(n_eq * droplets)]
        py = [O.Stream.replay(u_py[base + i]) for i in range(n_eq * droplets)]
        want, wN, wm = O.ptxc(1, g, L, O.all_classes(g, L, qs[s]), 0.1, 0.25, droplets, Nc, steps, nb, py, iters=iters,
                              want_hist=True)
        assert np.array_equal(mh[s], wm), "m(n) differs"
        assert np.array_equal(Nh[s], wN), "N(n) differs"
        np.testing.assert_allclose(out[s], want, rtol=1e-9)


def test_ptrc_replay_matches_reference_golden(ctx):
    n = 0
    for c in golden("shipped"):
        if c["kind"] != "ptrc":
            continue
        g, L, Nc = O.GEOM[c["geom"]], c["L"], c["Nc"]
        n_eq, steps = O.neq(g), c["steps"] // c["Nc"]
        u_nb, u_py = _golden_ladder_class_streams(c, n_eq, Nc, steps)
        out, _ = ctx.ptrc(g, L, c["inits"].reshape(1, n_eq, -1).copy(), c["p_error"], c["p_sampling"], 1, Nc, steps, per_class=True,
                          u_nb=u_nb, u_py=u_py)
        assert np.array_equal(out[0].astype(np.uint8), c["out"]) or np.array_equal(np.floor(out[0] + 1e-9).astype(np.uint8), c["out"])
        n += 1
    assert n >= 2


# ------------------------------------------------------------------ replay: PTEQ_alpha_with_shortest
@pytest.mark.parametrize("kind,g,L,bottom,b", [(1, O.PLANAR, 5, 0.15, 2.0), (1, O.XZZX, 5, 0.2, 1.5), (1, O.TORIC, 5, 0.15, 2.0),
                                               (0, O.ROTATED, 5, 0.1, 0.0), (2, O.XZZX, 7, 0.1, 6.0)])
def test_pteq_shortest_replay_matches_oracle(ctx, kind, g, L, bottom, b):
    rng = np.random.default_rng(800 + 10 * kind + g + L)
    S, iters, cap = 3, 10, 2500
    qs = [rand_lattice(rng, g, L, 0.1) for _ in range(S)]
    qm = np.stack([q.reshape(-1) for q in qs])
    n_nb, n_py = _ladder_budget(g, L, iters, cap)
    u_nb, u_py = rng.random((S, n_nb)), rng.random((S, n_py))
    pct, slen, sn, su, info = ctx.pteq_shortest(g, L, kind, qm, bottom, param_b=b, steps=cap, iters=iters, u_nb=u_nb, u_py=u_py)
    for s in range(S):
        want, _, _, winfo = O.pteq_with_shortest(kind, g, L, qs[s], bottom, O.Stream.replay(u_nb[s]), O.Stream.replay(u_py[s]),
                                                 param_b=b, steps=cap, iters=iters)
        assert info["steps"][s] == winfo["steps"]
        assert np.array_equal(pct[s], want)
        assert np.array_equal(slen[s], winfo["short_len"]), (slen[s], winfo["short_len"])
        assert np.array_equal(sn[s], winfo["short_n"])
        assert np.array_equal(su[s], winfo["short_unique"]), (su[s], winfo["short_unique"])


def test_pteq_shortest_replay_matches_reference_golden(ctx):
    n = 0
    for c in golden("shipped"):
        if c["kind"] != "pteq_alpha_shortest":
            continue
        g, L = O.GEOM[c["geom"]], c["L"]
        _, _, _, winfo = O.pteq_with_shortest(1, g, L, c["q"], c["p"], O.Stream.mt(c["nb_seed"]), O.Stream.py(c["py_seed"]),
                                              param_b=c["b"], steps=c["steps"])
        used = int(winfo["steps"])
        n_nb, n_py = _ladder_budget(g, L, 10, used)
        u_nb = np.random.RandomState(c["nb_seed"]).random_sample(n_nb).reshape(1, -1)
        u_py = _py_uniforms(c["py_seed"], n_py).reshape(1, -1)
        pct, slen, sn, su, info = ctx.pteq_shortest(g, L, 1, c["q"].reshape(1, -1).copy(), c["p"], param_b=c["b"], steps=used,
                                                    u_nb=u_nb, u_py=u_py)
        assert np.array_equal(pct[0], c["out"])
        z = su[0] * np.exp(np.log(c["p"]) * slen[0])          # decoders_biasednoise.py:163-169
        np.testing.assert_allclose(z / z.sum() * 100, c["out_unique"], rtol=1e-9)
        np.testing.assert_allclose(sn[0] / sn[0].sum() * 100, c["out_shortest_n"], rtol=1e-12)
        n += 1
    assert n >= 3


# ------------------------------------------------------------------ replay: general (x, y, z) noise
@pytest.mark.parametrize("gcode,gchain,L,droplets,xyz", [(O.PLANAR, O.PLANAR, 5, 1, True), (O.PLANAR, O.PLANAR, 7, 3, True),
                                                         (O.PLANAR, O.PLANAR, 5, 2, False), (O.TORIC, O.TORIC, 5, 2, True),
                                                         (O.TORIC, O.PLANAR, 5, 1, False)])
def test_stdc_general_noise_replay_matches_oracle(ctx, gcode, gchain, L, droplets, xyz):
    rng = np.random.default_rng(8800 + gcode + L + droplets)
    S, steps, iters = 2, 150, 5
    n_eq = O.neq(gcode)
    p_xyz = np.array([0.03, 0.02, 0.07])
    ps = np.array([0.09, 0.05, 0.12]) if xyz else float(p_xyz.sum())
    qs = [rand_lattice(rng, gcode, L, 0.1) for _ in range(S)]
    qm = np.stack([O.all_classes(gcode, L, q) for q in qs])
    u_nb = rng.random((S * n_eq * droplets, steps * iters, 4))
    out, out_s, distinct, st = ctx.stdc_general_noise(gcode, gchain, L, qm, p_xyz, ps, droplets, steps, iters=iters, per_class=True,
                                                      u_nb=u_nb)
    for s in range(S):
        base = s * n_eq * droplets
        nb = [O.Stream.replay(u_nb[base + i].reshape(-1)) for i in range(n_eq * droplets)]
        want, want_s, wd = O.stdc_general_noise(gcode, gchain, L, qm[s], p_xyz, ps, droplets, steps, nb, iters=iters)
        assert np.array_equal(distinct[s], wd)
        np.testing.assert_allclose(out[s], want, rtol=1e-9)
        np.testing.assert_allclose(out_s[s], want_s, rtol=1e-9)


def test_stdc_general_noise_replay_matches_reference_golden(ctx):
    n = 0
    for c in golden("shipped"):
        if c["kind"] != "stdc_general_noise":
            continue
        g, cg, L, steps = O.GEOM[c["geom"]], O.GEOM[c["chain_geom"]], c["L"], c["steps"]
        n_eq = O.neq(g)
        ps = c["p_sampling"] if c["p_sampling"].size else float(c["p_xyz"].sum())
        u_nb = np.random.RandomState(c["nb_seed"]).random_sample(n_eq * steps * 5 * 4).reshape(n_eq, steps * 5, 4)
        out, out_s, _, _ = ctx.stdc_general_noise(g, cg, L, c["inits"].reshape(1, n_eq, -1).copy(), c["p_xyz"], ps, 1, steps,
                                                  per_class=True, u_nb=u_nb)
        np.testing.assert_allclose(out[0], c["out"], rtol=1e-9)
        np.testing.assert_allclose(out_s[0], c["out_shortest"], rtol=1e-9)
        n += 1
    assert n >= 2


def test_chain_xyz_replay_matches_reference_golden_and_oracle(ctx):
    """Chain_xyz.update_chain_fast: the reference's own trajectory from its numba draws, and the oracle on random draws."""
    n = 0
    for c in golden("shipped"):
        if c["kind"] != "chain_fast_xyz":
            continue
        g, L, iters, blocks = O.GEOM[c["chain_geom"]], c["L"], c["iters"], c["blocks"]
        u = np.random.RandomState(c["nb_seed"]).random_sample(iters * blocks * 4).reshape(blocks, iters, 4)
        cur = c["q"].reshape(1, -1).copy()
        for b in range(blocks):
            ctx.chain_update_xyz(g, L, cur, c["p_sampling"], iters, u=u[b:b + 1])
            assert np.array_equal(cur[0], c["out"][b].reshape(-1)), b
        n += 1
    assert n >= 1
    rng = np.random.default_rng(99)
    for g, L in ((O.TORIC, 7), (O.PLANAR, 19), (O.ROTATED, 9), (O.XZZX, 5)):
        k = 3 if g in (O.TORIC, O.PLANAR) else 5
        ps = np.array([0.07, 0.11, 0.05])
        qm = np.stack([rand_lattice(rng, g, L, 0.15).reshape(-1) for _ in range(20)])
        u = rng.random((20, 200, k + 1))
        got = qm.copy()
        ctx.chain_update_xyz(g, L, got, ps, 200, u=u)
        for i in range(20):
            want = O.update_chain_fast_xyz(g, L, qm[i], ps / (1 - ps.sum()), 200, O.Stream.replay(u[i].reshape(-1)))
            assert np.array_equal(got[i], want)


# ------------------------------------------------------------------ key logs + dedupe kernel vs the HBM set
@pytest.mark.parametrize("g,L,droplets,steps", [(O.TORIC, 7, 16, 20000), (O.PLANAR, 9, 64, 6000), (O.TORIC, 15, 7, 9001),
                                                (O.TORIC, 5, 1, 333), (O.TORIC, 9, 64, 12000), (O.TORIC, 15, 32, 4000),
                                                (O.PLANAR, 7, 16, 2401), (O.TORIC, 5, 2, 3000), (O.TORIC, 7, 128, 1500),
                                                (O.PLANAR, 15, 16, 5000), (O.TORIC, 9, 10, 8000), (O.PLANAR, 11, 3, 7000),
                                                (O.TORIC, 5, 100, 900), (O.TORIC, 7, 600, 500)])
def test_log_dedupe_equals_table_set(ctx, g, L, droplets, steps, monkeypatch):
    """The same native chains counted in every way the library has: bucket logs written by the chain kernel + one-pass
    dedupe (mode 6, the default; droplets = 7, 10, 3, 100, 600 give CTAs that are not a multiple of 32 threads and a
    partly filled last CTA), per-chain key logs reduced by
    log_dedupe_kernel (mode 4), and the open-addressing set in HBM (modes 2 and 0).  N(n), the class distributions'
    inputs, must agree exactly -- this exercises the multi-bucket paths (tens of thousands of keys per table) that the
    small replay cases do not."""
    rng = np.random.default_rng(4000 + L)
    S = 5
    qm = np.stack([rand_lattice(rng, g, L, 0.12).reshape(-1) for _ in range(S)])
    outs = {}
    for mode in ("default", "4", "2", "0"):
        ctx.debug_set("insert_mode", -1 if mode == "default" else int(mode))
        try:
            out, st, hist = ctx.stdc(g, g, L, qm, 0.12, 0.3, droplets, steps, seed=99, want_hist=True)
        finally:
            ctx.debug_set("insert_mode", -1)
        outs[mode] = (out, st, hist)
    assert outs["4"][1]["table_slots"] == 0 and outs["2"][1]["table_slots"] > 0      # really different paths
    assert outs["default"][1]["table_slots"] == -1   # bucket logs for any droplets <= 1024 (CTAs of whole tables)
    for mode in ("4", "2", "0"):
        assert np.array_equal(outs["default"][2], outs[mode][2])
        assert outs["default"][1]["distinct"] == outs[mode][1]["distinct"]
        assert np.allclose(outs["default"][0], outs[mode][0], rtol=1e-12)
    assert outs["4"][1]["distinct"] > 0.1 * S * O.neq(g) * droplets * steps * 0.1


@pytest.mark.parametrize("droplets", [10, 64])
def test_waves_under_a_small_table_budget_give_the_same_result(droplets):
    """A table budget that holds only part of the batch makes the call run in waves (rounded to whole rounds of CTAs over
    the SMs); chains are seeded by their global index, so the result must not depend on the split."""
    from mcmc_qec_toric_rl_b200 import _lib
    g, L, steps, S = O.TORIC, 5, 2000, 700
    rng = np.random.default_rng(31)
    qm = np.stack([rand_lattice(rng, g, L, 0.1).reshape(-1) for _ in range(S)])
    c1 = _lib.Context(0)
    a = c1.stdc(g, g, L, qm, 0.1, 0.25, droplets, steps, seed=8, want_hist=True)
    c2 = _lib.Context(0)
    max_keys = droplets * steps
    nbc = 1
    while nbc < 128 and nbc * 20000 < max_keys:
        nbc *= 2
    per_syndrome = 16 * (nbc * ((max_keys + nbc - 1) // nbc + 65) + max(1024, max_keys // 16)) * 8
    c2.set_table_budget(per_syndrome * 333)
    b = c2.stdc(g, g, L, qm, 0.1, 0.25, droplets, steps, seed=8, want_hist=True)
    assert a[1]["waves"] == 1 and b[1]["waves"] >= 3 and b[1]["table_slots"] == -1
    assert np.array_equal(a[2], b[2]) and a[1]["distinct"] == b[1]["distinct"]
    assert np.allclose(a[0], b[0], rtol=1e-12)


@pytest.mark.parametrize("g,L", [(O.TORIC, 9), (O.PLANAR, 11), (O.PLANAR, 16)])
def test_row_word_width_does_not_change_native_results(ctx, g, L, monkeypatch):
    """debug_set("force_wide") runs a lattice with L <= 16 through the 64-bit row-word kernels (the ones L > 16 uses): the
    same seeds must give the same chains, i.e. identical N(n) and distinct counts, as the 32-bit kernels."""
    rng = np.random.default_rng(640 + L)
    qm = np.stack([rand_lattice(rng, g, L, 0.12).reshape(-1) for _ in range(6)])
    a = ctx.stdc(g, g, L, qm, 0.12, 0.3, 16, 4000, seed=21, want_hist=True)
    ctx.debug_set("force_wide", 1)
    try:
        b = ctx.stdc(g, g, L, qm, 0.12, 0.3, 16, 4000, seed=21, want_hist=True)
    finally:
        ctx.debug_set("force_wide", -1)
    assert b[1]["table_slots"] == 0 and a[1]["table_slots"] == -1       # per-chain logs behind the wide kernels, bucket logs otherwise
    assert np.array_equal(a[2], b[2]) and a[1]["distinct"] == b[1]["distinct"] and a[1]["accepted"] == b[1]["accepted"]


def test_more_droplets_than_a_cta_holds_use_per_chain_logs(ctx, monkeypatch):
    """droplets > 1024: a table's chains no longer fit one CTA, so the call takes per-chain logs; same counts as the HBM set."""
    g, L, droplets, steps = O.TORIC, 5, 1100, 300
    rng = np.random.default_rng(78)
    qm = np.stack([rand_lattice(rng, g, L, 0.1).reshape(-1) for _ in range(2)])
    a = ctx.stdc(g, g, L, qm, 0.1, 0.3, droplets, steps, seed=5, want_hist=True)
    ctx.debug_set("insert_mode", 2)
    try:
        b = ctx.stdc(g, g, L, qm, 0.1, 0.3, droplets, steps, seed=5, want_hist=True)
    finally:
        ctx.debug_set("insert_mode", -1)
    assert a[1]["table_slots"] == 0 and b[1]["table_slots"] > 0
    assert np.array_equal(a[2], b[2]) and a[1]["distinct"] == b[1]["distinct"]


def test_bucket_log_overflow_falls_back(ctx):
    """Chains at a high sampling rate offer a key at almost every sample; tiny tables make the fixed-capacity bucket logs
    and their overflow area run over, and the call must then redo itself with per-chain logs and still be exact."""
    g, L, droplets, steps = O.TORIC, 5, 64, 3000
    rng = np.random.default_rng(77)
    qm = np.stack([rand_lattice(rng, g, L, 0.3).reshape(-1) for _ in range(3)])
    a = ctx.stdc(g, g, L, qm, 0.3, 0.45, droplets, steps, seed=5, want_hist=True)
    ctx.debug_set("insert_mode", 2)
    try:
        b = ctx.stdc(g, g, L, qm, 0.3, 0.45, droplets, steps, seed=5, want_hist=True)
    finally:
        ctx.debug_set("insert_mode", -1)
    assert np.array_equal(a[2], b[2]) and a[1]["distinct"] == b[1]["distinct"]


# ------------------------------------------------------------------ PTDC with the conv_mult early stop
@pytest.mark.parametrize("g,L,Nc,droplets,per_class,conv", [(O.TORIC, 5, 4, 1, False, 2.0), (O.TORIC, 5, 3, 2, False, 1.5),
                                                             (O.PLANAR, 5, 4, 3, True, 2.0), (O.ROTATED, 5, 4, 2, False, 1.2)])
def test_ptdc_early_stop_replay_matches_oracle(ctx, g, L, Nc, droplets, per_class, conv):
    """PTDC_droplet's early stop (decoders.py:156-161): "new" is new to the droplet (a set per ladder on the device, united
    per class afterwards).  Every droplet must stop after the same number of Ladder.step calls as the oracle's, and the
    class distributions must agree."""
    rng = np.random.default_rng(7300 + g + L + Nc)
    S, steps, iters = 2, 400, 10
    n_eq = O.neq(g)
    qs = [rand_lattice(rng, g, L, 0.08) for _ in range(S)]
    qm = np.stack([O.all_classes(g, L, q) for q in qs]) if per_class else np.stack([q.reshape(-1) for q in qs])
    u_nb, u_py = _ptxc_streams(rng, g, S * n_eq * droplets, Nc, iters, steps)
    out, st, done = ctx.ptdc(g, L, qm, 0.1, 0.25, droplets, Nc, steps, iters=iters, per_class=per_class, u_nb=u_nb, u_py=u_py,
                             conv_mult=conv, want_steps=True)
    stopped = 0
    for s in range(S):
        base = s * n_eq * droplets
        nb = [O.Stream.replay(u_nb[base + i]) for i in range(n_eq * droplets)]
        py = [O.Stream.replay(u_py[base + i]) for i in range(n_eq * droplets)]
        want, wdone, _ = O.ptdc_conv(g, L, O.all_classes(g, L, qs[s]), 0.1, 0.25, droplets, Nc, steps, conv, nb, py, iters=iters)
        assert np.array_equal(done[s], wdone), (done[s], wdone)
        np.testing.assert_allclose(out[s], want, rtol=1e-9)
        stopped += int((wdone < steps).sum())
    assert stopped > 0, "no droplet stopped early: the case does not exercise the rule"
