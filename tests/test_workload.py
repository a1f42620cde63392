"""The workload steps around the decoders (generate_data.py:53-261): generate_random_error, eq_true,
apply_random_logical, failure counting.  CPU tests pin the oracle to the reference's own seeded outputs
(tests/golden/golden_workload.npz, made by make_golden_workload.py); GPU tests run the CUDA kernels through the
C ABI on the same uniforms (bit-exact) and check the native Philox statistics."""
import os

import numpy as np
import pytest

from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))


def cases():
    z = np.load(os.path.join(HERE, "golden", "golden_workload.npz"))
    out = []
    for i in range(int(z["n_cases"])):
        k = f"c{i}_"
        out.append({f[len(k):]: z[f] for f in z.files if f.startswith(k)})
    return out


CASES = cases()


def _args(c):
    g, L = O.GEOM[str(c["geom"])], int(c["L"])
    p_error, px, py, pz = [float(x) for x in c["params"]]
    if g == O.TORIC:
        return g, L, dict(p_error=p_error, p_xyz=None, pauli=c["pauli"])
    return g, L, dict(p_error=None, p_xyz=(px, py, pz), pauli=None)


# ------------------------------------------------------------------ oracle vs the reference's vectors (CPU)
def test_golden_covers_every_code():
    assert len(CASES) == 48
    assert {str(c["geom"]) for c in CASES} == {"toric", "planar", "rotated", "xzzx"}


@pytest.mark.parametrize("i", range(len(CASES)))
def test_oracle_reproduces_reference_workload_steps(i):
    c = CASES[i]
    g, L, kw = _args(c)
    q = O.generate_errors(g, L, c["u"], **kw)
    assert np.array_equal(q, c["q"])
    assert O.eq_class(g, L, q) == int(c["eq_true"])
    q2, _ = O.apply_random_logical(g, L, q, O.Stream.replay(c["u_log"]))
    assert np.array_equal(q2, c["q_hidden"])
    assert O.eq_class(g, L, q2) == int(c["eq_hidden"])


def test_oracle_count_failures_matches_numpy():
    rng = np.random.default_rng(5)
    d = rng.random((200, 16))
    d[3, 4] = d[3, 9] = 2.0            # tie: first maximum wins
    d[7, 5] = np.nan                   # numpy: a NaN is the arg-extremum
    truth = rng.integers(0, 16, 200).astype(np.int32)
    for use_min in (False, True):
        want = (np.argmin(d, 1) if use_min else np.argmax(d, 1))
        fails, choice = O.count_failures(d, truth, use_min)
        assert np.array_equal(choice, want)
        assert fails == int((want != truth).sum())


# ------------------------------------------------------------------ CUDA kernels (C ABI)
@pytest.fixture(scope="module")
def ctx():
    from mcmc_qec_toric_rl_b200 import _lib
    return _lib.default_context(0)


@pytest.mark.gpu
def test_generate_errors_replay_matches_reference(ctx):
    """every golden case, batched per (geom, L): lattices and eq_true bit-exact"""
    groups = {}
    for c in CASES:
        groups.setdefault((str(c["geom"]), int(c["L"])), []).append(c)
    for (gname, L), cs in groups.items():
        g = O.GEOM[gname]
        # one call per case: each case has its own probabilities
        for c in cs:
            _, _, kw = _args(c)
            qm, cls = ctx.generate_errors(g, L, 1, p_error=kw["p_error"], p_xyz=kw["p_xyz"], u=c["u"], pauli=kw["pauli"])
            assert np.array_equal(qm[0], c["q"]), (gname, L)
            assert int(cls[0]) == int(c["eq_true"])


@pytest.mark.gpu
@pytest.mark.parametrize("g,L", [(O.TORIC, 5), (O.TORIC, 15), (O.PLANAR, 7), (O.PLANAR, 21), (O.ROTATED, 25), (O.XZZX, 21),
                                 (O.TORIC, 32), (O.ROTATED, 3)])
def test_generate_errors_replay_matches_oracle(ctx, g, L):
    rng = np.random.default_rng(77 + 5 * g + L)
    S, n = 37, O.nsites(g, L)
    u = rng.random((S, n))
    if g == O.TORIC:
        pa = rng.integers(1, 4, (S, n)).astype(np.uint8)
        qm, cls = ctx.generate_errors(g, L, S, p_error=0.17, u=u, pauli=pa)
        want = np.stack([O.generate_errors(g, L, u[s], p_error=0.17, pauli=pa[s]) for s in range(S)])
    else:
        p = (0.04, 0.02, 0.11)
        qm, cls = ctx.generate_errors(g, L, S, p_xyz=p, u=u)
        want = np.stack([O.generate_errors(g, L, u[s], p_xyz=p) for s in range(S)])
    assert np.array_equal(qm, want)
    assert np.array_equal(cls, [O.eq_class(g, L, want[s]) for s in range(S)])
    assert np.array_equal(ctx.define_equivalence_class(g, L, want), cls)


@pytest.mark.gpu
def test_random_logical_replay_matches_reference(ctx):
    for c in CASES:
        g, L = O.GEOM[str(c["geom"])], int(c["L"])
        q2, ops = ctx.apply_random_logical(g, L, c["q"][None, :], u=c["u_log"][None, :])
        assert np.array_equal(q2[0], c["q_hidden"]), (str(c["geom"]), L)
        assert int(ctx.define_equivalence_class(g, L, q2)[0]) == int(c["eq_hidden"])


@pytest.mark.gpu
@pytest.mark.parametrize("g,L", [(O.TORIC, 15), (O.PLANAR, 11), (O.ROTATED, 25), (O.XZZX, 21)])
def test_random_logical_replay_matches_oracle(ctx, g, L):
    rng = np.random.default_rng(300 + g)
    S, n = 64, O.nsites(g, L)
    q = rng.integers(0, 4, (S, n)).astype(np.uint8)
    u = rng.random((S, 6))
    q2, ops = ctx.apply_random_logical(g, L, q, u=u)
    for s in range(S):
        want, _ = O.apply_random_logical(g, L, q[s], O.Stream.replay(u[s]))
        assert np.array_equal(q2[s], want)
    # native draws: syndrome-preserving, every operator reachable, class moves as the operator says
    q3, ops3 = ctx.apply_random_logical(g, L, np.repeat(q[:1], 4096, 0), seed=11)
    cls = ctx.define_equivalence_class(g, L, q3)
    n_ops = 16 if g == O.TORIC else 4
    counts = np.bincount(ops3[:, 0] + 4 * ops3[:, 1], minlength=n_ops)
    assert counts.min() > 4096 / n_ops * 0.6          # uniform over operators (binomial, > 8 sigma slack)
    assert len(np.unique(cls)) == n_ops               # each operator lands in its own class


@pytest.mark.gpu
@pytest.mark.parametrize("g", [O.TORIC, O.PLANAR, O.ROTATED, O.XZZX])
def test_generate_errors_native_statistics(ctx, g):
    """Philox draws: per-Pauli rates within 5 sigma of the requested probabilities; planar's unused sites stay 0"""
    L, S = 9, 4000
    n = O.nsites(g, L)
    if g == O.TORIC:
        qm, cls = ctx.generate_errors(g, L, S, p_error=0.15, seed=5)
        want = {1: 0.05, 2: 0.05, 3: 0.05}
    else:
        qm, cls = ctx.generate_errors(g, L, S, p_xyz=(0.03, 0.01, 0.12), seed=5)
        want = {1: 0.03, 2: 0.01, 3: 0.12}
    live = np.ones(n, bool)
    if g == O.PLANAR:
        m = np.ones((2, L, L), bool)
        m[1, -1, :] = False
        m[1, :, -1] = False
        live = m.reshape(-1)
        assert not qm[:, ~live].any()
    N = S * live.sum()
    for v, p in want.items():
        k = (qm[:, live] == v).sum()
        assert abs(k - N * p) < 5 * np.sqrt(N * p * (1 - p)), (v, k / N, p)
    assert np.array_equal(cls, [O.eq_class(g, L, qm[s]) for s in range(S)])
    qm2, _ = ctx.generate_errors(g, L, S, p_error=0.15 if g == O.TORIC else None, p_xyz=None if g == O.TORIC else (0.03, 0.01, 0.12),
                                 seed=6)
    assert not np.array_equal(qm, qm2)                # the seed matters
    # sites are independent: neighbouring-site correlation of the error indicator ~ 0
    e = (qm[:, live] != 0).astype(np.float64)
    cc = np.corrcoef(e[:, :-1].reshape(-1), e[:, 1:].reshape(-1))[0, 1]
    assert abs(cc) < 0.01


@pytest.mark.gpu
def test_count_failures_matches_numpy(ctx):
    rng = np.random.default_rng(9)
    d = rng.random((1000, 16))
    d[3, 4] = d[3, 9] = 2.0
    d[7, 5] = np.nan
    truth = rng.integers(0, 16, 1000).astype(np.int32)
    for use_min in (False, True):
        want = np.argmin(d, 1) if use_min else np.argmax(d, 1)
        fails, choice = ctx.count_failures(d, truth, use_min)
        assert np.array_equal(choice, want)
        assert fails == int((want != truth).sum())
    d8 = rng.integers(0, 101, (513, 4)).astype(np.uint8)
    t8 = rng.integers(0, 4, 513).astype(np.int32)
    fails, choice = ctx.count_failures(d8, t8)
    assert np.array_equal(choice, np.argmax(d8, 1))
    assert fails == int((np.argmax(d8, 1) != t8).sum())


# ------------------------------------------------------------------ the whole loop (generate_data.py:19-261)
@pytest.mark.gpu
def test_generate_batch_device_resident_stdc(ctx):
    """toric d=5, STDC: errors, eq_true, hidden class, distributions and the failure count all come from one pass
    on the device; the count equals argmax != eq_true recomputed on the host, the syndrome of each data point is the
    syndrome of its recorded error (the hiding operator is a logical), and decoding succeeds at the usual rate."""
    from mcmc_qec_toric_rl_b200 import generate_data as G
    params = dict(code='toric', method='STDC', size=5, noise='depolarizing', p_error=0.08, p_sampling=0.25, droplets=8,
                  steps=2000, mwpm_init=False)
    res = G.generate_batch(params, 200, seed=3)
    assert res['qubit'].shape == (200, 50) and res['distr'].shape == (200, 16)
    assert np.array_equal(res['eq_true'], [O.eq_class(O.TORIC, 5, q) for q in res['qubit']])
    assert np.allclose(res['distr'].sum(1), 100.0)
    assert np.array_equal(res['choice'], res['distr'].argmax(1))
    assert res['failures'] == int((res['distr'].argmax(1) != res['eq_true']).sum())
    assert res['failures'] <= 30                       # ~5 % logical failures at p = 0.08, d = 5
    rate = (res['qubit'] != 0).mean()
    assert abs(rate - 0.08) < 5 * np.sqrt(0.08 * 0.92 / res['qubit'].size)


@pytest.mark.gpu
@pytest.mark.parametrize("params", [
    dict(code='planar', method='STDC', size=5, noise='depolarizing', p_error=0.06, p_sampling=0.25, droplets=4, steps=1500),
    dict(code='toric', method='ST', size=5, noise='depolarizing', p_error=0.05, steps=1500),
    dict(code='toric', method='STRC', size=5, noise='depolarizing', p_error=0.06, p_sampling=0.25, droplets=4, steps=1500),
    dict(code='rotated', method='PTEQ', size=5, noise='depolarizing', p_error=0.06, pt_steps=40000),
    dict(code='xzzx', method='PTEQ', size=5, noise='biased', p_error=0.06, eta=10, pt_steps=40000),
    dict(code='xzzx', method='PTEQ', size=5, noise='alpha', p_error=0.05, alpha=2, pt_steps=40000),
    dict(code='planar', method='STDC_N_n', size=5, noise='alpha', p_error=0.05, p_sampling=0.3, alpha=2, steps=1500),
    dict(code='planar', method='PTDC', size=5, noise='depolarizing', p_error=0.06, p_sampling=0.25),
])
def test_generate_batch_every_method(ctx, params):
    from mcmc_qec_toric_rl_b200 import generate_data as G
    params = dict(params, mwpm_init=False)
    S = 48
    res = G.generate_batch(params, S, seed=21)
    n_eq = 16 if params['code'] == 'toric' else 4
    assert res['distr'].shape[0] == S and res['distr'].shape[1] >= n_eq
    pick = res['distr'][:, :n_eq].argmin(1) if params['method'] == 'ST' else res['distr'][:, :n_eq].argmax(1)
    assert np.array_equal(res['choice'], pick)
    assert res['failures'] == int((pick != res['eq_true']).sum())
    if params['method'] != 'STDC_N_n':                 # the reference does not score STDC_N_n either
        assert res['failures'] <= S * 0.45, res['failures']


@pytest.mark.gpu
def test_generate_writes_the_reference_file_format(ctx, tmp_path):
    import pandas as pd
    from mcmc_qec_toric_rl_b200 import generate_data as G
    from mcmc_qec_toric_rl_b200.src.mcmc import MCMCDataReader
    params = dict(code='toric', method='STDC', size=5, noise='depolarizing', p_error=0.1, p_sampling=0.25, droplets=4,
                  steps=500, mwpm_init=False)
    path = str(tmp_path / 'data.xz')
    failed, made = G.generate(path, params, nbr_datapoints=70, batch=32, seed=4, verbose=False)
    assert made == 70
    df = pd.read_pickle(path)
    assert list(df.index.names) == ['data_nr', 'type'] and list(df.columns) == ['data']
    assert df.loc[(-1, 0), 'data'] == params
    assert len(df) == 1 + 2 * 70
    q = df.loc[(13, 0), 'data']
    d = df.loc[(13, 1), 'data']
    assert q.shape == (2, 5, 5) and q.dtype == np.uint8 and d.shape == (16,)
    rd = MCMCDataReader(path, 5)
    assert rd.get_capacity() == 70 and rd.has_next() and len(rd.full()) == 1 + 2 * 70
    # fixed_errors stops at the requested number of failures
    failed2, made2 = G.generate(path, dict(params, p_error=0.2), nbr_datapoints=5, fixed_errors=3, batch=16, seed=5, verbose=False)
    assert failed2 == 3 and made2 <= 10000000
    df2 = pd.read_pickle(path)
    assert len(df2) == 1 + 2 * made2
    # batch=None: the loop sizes its batches from the library's plan (a first batch of 256, then whole rounds of CTAs)
    planar = dict(code='planar', method='STDC', size=5, noise='depolarizing', p_error=0.1, p_sampling=0.25, droplets=4,
                  steps=200, mwpm_init=False)
    failed3, made3 = G.generate(path, planar, nbr_datapoints=3000, seed=6, verbose=False)
    assert made3 == 3000 and len(pd.read_pickle(path)) == 1 + 2 * 3000
    sms = ctx.device_info()["sm_count"]
    assert G.auto_batch(planar) % (sms * 1024 // (4 * 4)) == 0          # whole rounds of 1024-thread CTAs, 4 classes x 4 chains
    assert G.auto_batch(dict(planar, method='PTEQ')) == 32 * sms


# ------------------------------------------------------------------ PTDC with the conv_mult early stop
def ptdc_conv_cases():
    z = np.load(os.path.join(HERE, "golden", "golden_workload.npz"))
    out = []
    for i in range(int(z["n_ptdc_conv"])):
        k = f"p{i}_"
        out.append({f[len(k):]: z[f] for f in z.files if f.startswith(k)})
    return out


PTDC_CONV = ptdc_conv_cases()


@pytest.mark.parametrize("i", range(len(PTDC_CONV)))
def test_oracle_reproduces_reference_ptdc_early_stop(i):
    """PTDC(..., droplets=1, Nc=4, steps=2400, conv_mult) of the seeded reference (decoders.py:138-233)"""
    c = PTDC_CONV[i]
    g = O.GEOM[str(c["geom"])]
    n_eq = O.neq(g)
    nb, py = O.Stream.mt(int(c["seeds"][1])), O.Stream.py(int(c["seeds"][0]))
    out, done, _ = O.ptdc_conv(g, 5, c["inits"], 0.1, 0.25, 1, 4, 2400 // 4, float(c["conv_mult"]), [nb] * n_eq, [py] * n_eq)
    assert (done < 600).any(), "no ladder stopped early: the vector does not exercise the rule"
    # the reference truncates to uint8: allow the float result to sit within rounding of a truncation boundary
    assert np.array_equal(np.floor(out + 1e-9).astype(np.uint8), c["out"]) or np.array_equal(out.astype(np.uint8), c["out"]), (out, c["out"])


@pytest.mark.gpu
def test_threshold_sweep_through_the_sharded_driver(ctx):
    """Config 5's driver on one GPU (world size 1; the two-rank control flow is tests/test_sharding.py): every (d, p) point
    gets its syndromes through generate_batch, the gathered failure counts equal argmax != eq_true recomputed from the
    gathered distributions, and the failure rate rises with p at fixed d."""
    from mcmc_qec_toric_rl_b200 import generate_data as G, sharding
    points = [dict(d=d, p=p) for d in (5, 7) for p in (0.04, 0.12, 0.20)]
    truth = {}

    def decode_chunk(pt, n, item):
        params = dict(code='planar', method='STDC', size=pt['d'], noise='depolarizing', p_error=pt['p'], p_sampling=0.25,
                      droplets=4, steps=pt['d'] ** 4, mwpm_init=False)
        res = G.generate_batch(params, n, seed=1000 + item)
        truth[item] = res['eq_true']
        return res['failures'], np.concatenate([res['distr'], res['eq_true'][:, None].astype(np.float64)], axis=1)
    curve = sharding.run_sweep_sharded(points, 300, 128, decode_chunk, rank=0, world=1)
    assert [o['syndromes'] for o in curve] == [300] * 6
    for o in curve:
        distr, eq_true = o['extra'][:, :4], o['extra'][:, 4].astype(np.int64)
        assert o['failures'] == int((distr.argmax(1) != eq_true).sum())
        assert abs(o['rate'] - o['failures'] / 300) < 1e-12
    for d0 in (0, 3):
        rates = [curve[d0 + k]['rate'] for k in range(3)]
        assert rates[0] < rates[1] < rates[2], rates
        assert rates[0] < 0.1 and rates[2] > 0.15, rates
