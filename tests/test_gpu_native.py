"""GPU parity of the NATIVE (Philox) kernels against the oracle, bit for bit.

Philox4x32-10 is counter based and keyed by the chain's global index, so the exact words a device chain consumes can be
regenerated on the host: the oracle runs the reference's algorithm on those words (oracle/qec_oracle.c, "Native draws":
one word per stabilizer proposal over the canonical numbering, one bit per rain site, u = word / 2^32 for every accept
decision) and the device's distinct-chain histograms N(n), visit histograms m(n), accept counts and final lattices must
EQUAL the oracle's -- at the BASELINE sizes (toric d=15 with 64 chains per class in the 1024-thread bucket-log
instantiation, planar d=11, the 64-bit row-word kernels at d=21), not only at d=5.

The second half compares native decoding with the oracle's independent MT19937 runs on the same syndromes: logical
failure counts within the stated binomial confidence interval (north_star correctness (2))."""
import threading

import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu
RAIN_TAG = 0x80000000


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


def chain_streams(seed, s, n_eq, droplets):
    """Streams of the chains of syndrome s: chain id = (s * n_eq + eq) * droplets + d (qecmc_stdc_fast.cuh: gchain)."""
    ids = [(s * n_eq + eq) * droplets + d for eq in range(n_eq) for d in range(droplets)]
    return [O.Stream.philox(seed, i, 0) for i in ids], [O.Stream.philox(seed, i, RAIN_TAG) for i in ids]


def parallel(fn, items, threads=16):
    """run fn over items on host threads (the oracle's ctypes calls drop the GIL)"""
    out, lock, todo = {}, threading.Lock(), list(items)

    def work():
        while True:
            with lock:
                if not todo:
                    return
                it = todo.pop()
            out[it] = fn(it)

    ts = [threading.Thread(target=work) for _ in range(min(threads, len(todo)))]
    [t.start() for t in ts]
    [t.join() for t in ts]
    return out


def oracle_stdc(g, gc, L, qs, s, p_error, p_sampling, droplets, steps, seed, randomize=True, conv_mult=0.0):
    n_eq = O.neq(g)
    nb, np_ = chain_streams(seed, s, n_eq, droplets)
    return O.stdc(g, gc, L, O.all_classes(g, L, qs[s]), p_error, p_sampling, droplets, steps, nb, np_, randomize=randomize,
                  conv_mult=conv_mult, want_hist=True)


# (geometry, L, syndromes, droplets, samples, syndromes checked against the oracle)
STDC_CASES = [
    # the headline instantiation: toric d=15, 16 classes x 64 chains, one 1024-thread CTA per SM (148 syndromes fill the GPU)
    pytest.param(O.TORIC, 15, 148, 64, 2000, (0, 1, 63, 64, 100, 147), id="toric15-64chains-full-gpu"),
    pytest.param(O.TORIC, 15, 3, 64, 2500, (0, 1, 2), id="toric15-64chains-small-batch"),
    pytest.param(O.PLANAR, 11, 40, 64, 2000, (0, 17, 39), id="planar11-64chains"),
    pytest.param(O.PLANAR, 21, 6, 16, 2000, (0, 3, 5), id="planar21-packed-lattice"),
    pytest.param(O.TORIC, 17, 4, 8, 1500, (0, 3), id="toric17-packed-lattice"),
    pytest.param(O.PLANAR, 19, 3, 8, 1500, (0, 2), id="planar19-packed-lattice"),
    pytest.param(O.TORIC, 24, 2, 4, 1200, (0, 1), id="toric24-packed-lattice"),
    pytest.param(O.PLANAR, 25, 2, 4, 1000, (0, 1), id="planar25-u64-words"),
    pytest.param(O.TORIC, 5, 7, 10, 3125, (0, 3, 6), id="toric5-reference-defaults"),
    pytest.param(O.PLANAR, 15, 5, 16, 3000, (0, 4), id="planar15"),
    pytest.param(O.TORIC, 9, 4, 3, 4000, (0, 1, 2, 3), id="toric9-3chains"),
]


@pytest.mark.parametrize("g,L,S,droplets,steps,check", STDC_CASES)
def test_native_stdc_equals_oracle_on_the_same_words(ctx, g, L, S, droplets, steps, check):
    """STDC (decoders.py:236-322), native draws: N(n) per (syndrome, class) equal to the oracle's, distributions to 1e-9."""
    rng = np.random.default_rng(7000 + 31 * g + L)
    p_error, p_sampling, seed = 0.15 if L >= 11 else 0.1, 0.25, 0xC0FFEE + L
    qs = [rand_lattice(rng, g, L, p_error) for _ in range(S)]
    qm = np.stack([q.reshape(-1) for q in qs])
    out, st, hist = ctx.stdc(g, g, L, qm, p_error, p_sampling, droplets, steps, seed=seed, want_hist=True)
    assert st["metropolis_steps"] == S * O.neq(g) * droplets * steps * 5
    if g in (O.TORIC, O.PLANAR) and L <= 16:
        assert st["table_slots"] == -1          # bucket logs written by stdc_fast_kernel<.., BLOG = true>
    want = parallel(lambda s: oracle_stdc(g, g, L, qs, s, p_error, p_sampling, droplets, steps, seed), check)
    for s in check:
        w_out, w_distinct, w_hist = want[s]
        assert np.array_equal(hist[s].astype(np.int64), w_hist), f"N(n) differs from the oracle at syndrome {s}"
        # (L = 24 at p = 0.15: every exp(-beta n) underflows, the reference's 0 / 0 -- NaN on both sides)
        assert np.allclose(out[s], w_out, rtol=1e-9, atol=1e-300, equal_nan=True)
        assert w_distinct.sum() > O.neq(g) * droplets       # the chains moved


@pytest.mark.parametrize("g,L", [(O.PLANAR, 17), (O.PLANAR, 20), (O.PLANAR, 21), (O.PLANAR, 22), (O.TORIC, 18), (O.TORIC, 21),
                                 (O.TORIC, 23)])
def test_packed_lattice_kernel_equals_the_64_bit_row_word_kernel(ctx, g, L):
    """Two-layer codes with 17 <= L <= 24 run on the packed-lattice chain kernel (qecmc_stdc_pk.cuh); debug_set("packed", 0)
    keeps them on the 64-bit row-word kernel.  Same draws, same step: N(n), m(n), mean lengths and distributions are
    identical -- for every number of table copies, for per-chain logs and for the deferred HBM set, for a batch that fills
    several CTAs and for one that does not."""
    rng = np.random.default_rng(300 + L)
    S, droplets, steps, seed = (150 if L == 21 else 9), 16, 600, 9000 + L
    qm = np.stack([rand_lattice(rng, g, L, 0.14).reshape(-1) for _ in range(S)])

    def run_all():
        a = ctx.stdc(g, g, L, qm, 0.14, 0.25, droplets, steps, seed=seed, want_hist=True)
        b = ctx.strc(g, g, L, qm[:5], 0.14, 0.25, 6, steps, seed=seed + 1, want_hist=True)
        c = ctx.single_temp(g, g, L, qm[:7], 0.12, 700, seed=seed + 2)
        ctx.debug_set("insert_mode", 2)
        try:
            d = ctx.stdc(g, g, L, qm[:4], 0.14, 0.25, 5, steps, seed=seed + 3, want_hist=True)
        finally:
            ctx.debug_set("insert_mode", -1)
        return a, b, c, d

    if L == 18:
        # the reference as shipped proposes PLANAR stabilizers on the toric lattice (src/mcmc.py:6): boundary
        # stabilizers with missing slots on a lattice without unused sites
        run_toric = run_all

        def run_all():
            return run_toric() + (ctx.stdc(O.TORIC, O.PLANAR, L, qm[:6], 0.14, 0.25, 8, steps, seed=seed + 4, want_hist=True),)

    ctx.debug_set("packed", 0)
    try:
        ref = run_all()
    finally:
        ctx.debug_set("packed", -1)
    for rep in (-1, 2, 8):
        ctx.debug_set("packed", rep)
        try:
            got = run_all()
        finally:
            ctx.debug_set("packed", -1)
        same = lambda x, y: np.array_equal(x, y, equal_nan=True)    # a class without weight: 0 / 0 on both sides
        assert same(got[0][2], ref[0][2]) and same(got[0][0], ref[0][0])
        assert got[0][1]["accepted"] == ref[0][1]["accepted"] and got[0][1]["distinct"] == ref[0][1]["distinct"]
        assert same(got[1][2], ref[1][2]) and same(got[1][3], ref[1][3]) and same(got[1][0], ref[1][0])
        assert same(got[2][0], ref[2][0])
        assert same(got[3][2], ref[3][2]) and same(got[3][0], ref[3][0])
        if len(ref) > 4:
            assert same(got[4][2], ref[4][2]) and same(got[4][0], ref[4][0])


def test_last_plan_sizes_a_batch_that_fills_whole_rounds(ctx):
    """qecmc_last_plan: after a call, how many syndromes one wave may hold and how many chains one round of full-size CTAs
    holds -- a batch of that many syndromes runs as one wave, and the figures do not depend on the size of the call."""
    g, L, droplets, steps = O.PLANAR, 21, 16, 2000
    rng = np.random.default_rng(77)
    qm = np.stack([rand_lattice(rng, g, L, 0.12).reshape(-1) for _ in range(40)])
    ctx.stdc(g, g, L, qm[:3], 0.12, 0.25, droplets, steps, seed=3)
    wave_small, round_small = ctx.last_plan()
    _, st = ctx.stdc(g, g, L, qm, 0.12, 0.25, droplets, steps, seed=3)[:2]
    wave_cap, round_chains = ctx.last_plan()
    assert round_chains == round_small and abs(wave_cap - wave_small) <= 0.02 * wave_cap    # (free memory moves a little)
    assert wave_cap >= 40 and st["waves"] == 1
    sms = ctx.device_info()["sm_count"]
    assert round_chains % (sms * droplets) == 0                  # whole tables per CTA, one CTA per SM
    assert 16 * 32 * sms <= round_chains <= 1024 * sms            # the packed-lattice kernel keeps >= 16 warps per SM at d = 21
    g2, L2 = O.TORIC, 15
    q2 = np.stack([rand_lattice(rng, g2, L2, 0.12).reshape(-1) for _ in range(4)])
    ctx.stdc(g2, g2, L2, q2, 0.12, 0.25, 64, 500, seed=3)
    assert ctx.last_plan()[1] == 1024 * sms                       # the headline kernel: one 1024-thread CTA per SM


@pytest.mark.parametrize("mode", [4, 2, 0])
def test_native_stdc_other_insert_modes_equal_oracle(ctx, mode):
    """The per-chain-log, deferred and synchronous HBM-set paths against the oracle (not against each other)."""
    g, L, S, droplets, steps, seed = O.TORIC, 7, 5, 16, 3000, 77
    rng = np.random.default_rng(12)
    qs = [rand_lattice(rng, g, L, 0.1) for _ in range(S)]
    qm = np.stack([q.reshape(-1) for q in qs])
    ctx.debug_set("insert_mode", mode)
    try:
        out, st, hist = ctx.stdc(g, g, L, qm, 0.1, 0.25, droplets, steps, seed=seed, want_hist=True)
    finally:
        ctx.debug_set("insert_mode", -1)
    for s in (0, 4):
        w_out, _, w_hist = oracle_stdc(g, g, L, qs, s, 0.1, 0.25, droplets, steps, seed)
        assert np.array_equal(hist[s].astype(np.int64), w_hist)
        assert np.allclose(out[s], w_out, rtol=1e-9)


@pytest.mark.parametrize("g,L,droplets,conv", [(O.TORIC, 5, 1, 2.0), (O.TORIC, 7, 3, 1.5), (O.PLANAR, 7, 4, 2.0)])
def test_native_stdc_early_stop_equals_oracle(ctx, g, L, droplets, conv):
    """conv_mult != 0 (decoders.py:257-263) on native words: same stopping sample per chain, hence the same N(n) and
    the same number of Metropolis steps."""
    S, steps, seed = 3, 4000, 5150
    rng = np.random.default_rng(40 + L)
    qs = [rand_lattice(rng, g, L, 0.1) for _ in range(S)]
    qm = np.stack([q.reshape(-1) for q in qs])
    out, st, hist = ctx.stdc(g, g, L, qm, 0.1, 0.25, droplets, steps, seed=seed, conv_mult=conv, want_hist=True)
    assert st["metropolis_steps"] < S * O.neq(g) * droplets * steps * 5      # somebody stopped early
    for s in range(S):
        w_out, _, w_hist = oracle_stdc(g, g, L, qs, s, 0.1, 0.25, droplets, steps, seed, conv_mult=conv)
        assert np.array_equal(hist[s].astype(np.int64), w_hist)
        assert np.allclose(out[s], w_out, rtol=1e-9)


@pytest.mark.parametrize("g,L,S,droplets,steps", [(O.TORIC, 15, 148, 64, 1500), (O.PLANAR, 9, 6, 10, 3000), (O.PLANAR, 17, 4, 8, 2000)])
def test_native_strc_equals_oracle_on_the_same_words(ctx, g, L, S, droplets, steps):
    """STRC (decoders.py:745-949): m(n) and the (shortest, next shortest, distinct counts) bookkeeping, bit for bit."""
    rng = np.random.default_rng(8100 + L)
    p_error, seed = 0.12, 424242 + L
    qs = [rand_lattice(rng, g, L, p_error) for _ in range(S)]
    qm = np.stack([q.reshape(-1) for q in qs])
    out, st, mh, info = ctx.strc(g, g, L, qm, p_error, 0.25, droplets, steps, seed=seed, want_hist=True)
    n_eq = O.neq(g)
    check = sorted({0, S // 2, S - 1})

    def run(s):
        nb, np_ = chain_streams(seed, s, n_eq, droplets)
        return O.strc(g, g, L, O.all_classes(g, L, qs[s]), p_error, 0.25, droplets, steps, nb, np_, want_hist=True)

    want = parallel(run, check)
    for s in check:
        w_out, w_mh, w_info = want[s]
        assert np.array_equal(mh[s].astype(np.int64), w_mh), f"m(n) differs at syndrome {s}"
        assert np.array_equal(info[s].astype(np.int64), w_info)
        assert np.allclose(out[s], w_out, rtol=1e-9)


def test_native_single_temp_equals_oracle(ctx):
    """single_temp (decoders.py:108-135): the mean chain length per class from the same words."""
    g, L, S, max_iters, seed = O.TORIC, 9, 20, 3000, 99
    rng = np.random.default_rng(5)
    qs = [rand_lattice(rng, g, L, 0.1) for _ in range(S)]
    qm = np.stack([q.reshape(-1) for q in qs])
    out, st = ctx.single_temp(g, g, L, qm, 0.12, max_iters, seed=seed)
    for s in (0, 9, 19):
        nb = [O.Stream.philox(seed, s * 16 + eq, 0) for eq in range(16)]
        want = O.single_temp(g, g, L, O.all_classes(g, L, qs[s]), 0.12, max_iters, nb)
        assert np.array_equal(out[s], want)          # sums of integers divided once: exact


@pytest.mark.parametrize("g,L", [(O.TORIC, 15), (O.PLANAR, 21), (O.ROTATED, 25), (O.XZZX, 21), (O.ROTATED, 7)])
def test_native_chain_update_equals_oracle(ctx, g, L):
    """Chain.update_chain_fast (mcmc.py:45-46, 152-160) with native draws: final lattices and accept counts."""
    chains, iters, p, seed = 300, 5001, 0.2, 31337
    rng = np.random.default_rng(60 + L)
    q0 = np.stack([rand_lattice(rng, g, L, 0.12).reshape(-1) for _ in range(chains)])
    qm = q0.copy()
    st = ctx.chain_update(g, L, qm, p, iters, seed=seed)
    factor, acc = (p / 3.0) / (1.0 - p), 0
    for ch in (0, 1, 150, 299):
        want, dE, a = O.update_chain_fast(g, L, q0[ch], factor, iters, O.Stream.philox(seed, ch, 0), trace=True)
        assert np.array_equal(qm[ch], np.asarray(want).reshape(-1)), ch
    assert 0 < st["accepted"] < chains * iters
    # a second call continues the streams where the first stopped (stream_offset)
    qa, qb = q0[:4].copy(), q0[:4].copy()
    ctx.chain_update(g, L, qa, p, 1000, seed=seed)
    ctx.chain_update(g, L, qa, p, 1001, seed=seed, stream_offset=1000)
    ctx.chain_update(g, L, qb, p, 2001, seed=seed)
    assert np.array_equal(qa, qb)


@pytest.mark.parametrize("g,L,xyz", [(O.PLANAR, 7, True), (O.PLANAR, 9, False), (O.ROTATED, 7, True)])
def test_native_general_noise_equals_oracle(ctx, g, L, xyz):
    """STDC_general_noise(_shortest) (decoders.py:325-508) on native words: distinct counts and both distributions."""
    S, droplets, steps, seed = 3, 4, 3000, 2718
    rng = np.random.default_rng(3 + L)
    qs = [rand_lattice(rng, g, L, 0.1) for _ in range(S)]
    qm = np.stack([q.reshape(-1) for q in qs])
    p_xyz = np.array([0.02, 0.03, 0.07])
    p_sampling = np.array([0.06, 0.05, 0.09]) if xyz else 0.25
    out, out_s, distinct, st = ctx.stdc_general_noise(g, g, L, qm, p_xyz, p_sampling, droplets, steps, seed=seed)
    n_eq = O.neq(g)
    for s in range(S):
        nb = [O.Stream.philox(seed, (s * n_eq + eq) * droplets + d, 0) for eq in range(n_eq) for d in range(droplets)]
        w, w_s, w_d = O.stdc_general_noise(g, g, L, O.all_classes(g, L, qs[s]), p_xyz, p_sampling, droplets, steps, nb)
        assert np.array_equal(distinct[s], w_d)
        assert np.allclose(out[s], w, rtol=1e-9) and np.allclose(out_s[s], w_s, rtol=1e-9)


# ------------------------------------------------------------------ native vs the oracle's own MT19937 runs
def _binomial_gap_ok(f_a, f_b, n, z=3.0):
    """two failure counts out of n trials each (same syndromes, independent randomness): |difference| within z sigma of
    the pooled binomial, plus one count for discreteness"""
    p_hat = (f_a + f_b) / (2.0 * n)
    sigma = np.sqrt(2.0 * p_hat * (1.0 - p_hat) * n)
    return abs(int(f_a) - int(f_b)) <= z * sigma + 1, sigma


def test_toric_d15_failure_rate_within_binomial_ci_of_the_oracle(ctx):
    """BASELINE config 1 (toric d=15, p = 0.15, STDC, p_sampling 0.25), 200 syndromes: the native decode and the oracle's
    MT19937 decode of the SAME syndromes (droplets 16, 5000 samples per chain) agree on the logical failure count within 3
    sigma of the pooled binomial (stated CI).  At p = 0.15 the toric code is close to threshold and 5000 samples leave
    every STDC estimate noisy, so two runs of the ORACLE ITSELF with different seeds disagree on most syndromes; that
    oracle-vs-oracle scatter is the yardstick for the per-class distributions: the native-vs-oracle mean absolute
    difference may exceed it by at most 20 % + 0.5 points, and the share of syndromes on which both pick the same class
    may fall short of the oracle-vs-oracle share by at most 3 binomial sigma + 2 points."""
    g, L, S, droplets, steps = O.TORIC, 15, 200, 16, 5000
    rng = np.random.default_rng(20251)
    qs, truth = [], []
    for _ in range(S):
        q = rand_lattice(rng, g, L, 0.15)
        truth.append(O.eq_class(g, L, q))
        q2, _ = O.apply_random_logical(g, L, q, O.Stream.mt(int(rng.integers(1 << 30))))    # generate_data.py:131
        qs.append(np.asarray(q2, np.uint8).reshape(-1))
    qm, truth = np.stack(qs), np.array(truth)
    gpu, st = ctx.stdc(g, g, L, qm, 0.15, 0.25, droplets, steps, seed=5)
    ref = O.stdc_batch(g, g, L, qm, 0.15, 0.25, droplets, steps, seed=17, threads=16)
    ref2 = O.stdc_batch(g, g, L, qm, 0.15, 0.25, droplets, steps, seed=23, threads=16)
    f_gpu, f_ref, f_ref2 = (int((x.argmax(1) != truth).sum()) for x in (gpu, ref, ref2))
    ok, sigma = _binomial_gap_ok(f_gpu, f_ref, S)
    assert ok, (f_gpu, f_ref, sigma)
    assert _binomial_gap_ok(f_gpu, f_ref2, S)[0], (f_gpu, f_ref2)
    assert 0 < f_gpu < S and 0 < f_ref < S, (f_gpu, f_ref)
    d_go, d_oo = np.abs(gpu - ref).mean(), np.abs(ref - ref2).mean()
    assert d_go <= 1.2 * d_oo + 0.5, (d_go, d_oo)
    a_go, a_oo = (gpu.argmax(1) == ref.argmax(1)).mean(), (ref.argmax(1) == ref2.argmax(1)).mean()
    assert a_go >= a_oo - 3.0 * np.sqrt(max(a_oo * (1 - a_oo), 1e-9) / S) - 0.02, (a_go, a_oo)


# ------------------------------------------------------------------ tempering ladders on native words
# (kind, geometry, L, bottom, param_b, Nc, p_logical): 0 depolarizing, 1 alpha, 2 biased
LADDERS = [
    pytest.param(0, O.ROTATED, 9, 0.15, 0.0, 9, 0.5, id="rotated9-depol"),
    pytest.param(0, O.ROTATED, 25, 0.15, 0.0, 25, 0.5, id="rotated25-depol-u64"),
    pytest.param(0, O.TORIC, 5, 0.1, 0.0, 5, 0.5, id="toric5-depol"),
    pytest.param(0, O.TORIC, 9, 0.12, 0.0, 9, 0.3, id="toric9-depol"),
    pytest.param(0, O.PLANAR, 7, 0.12, 0.0, 3, 0.5, id="planar7-3rungs"),
    pytest.param(0, O.PLANAR, 17, 0.12, 0.0, 6, 0.5, id="planar17-u64"),
    pytest.param(0, O.ROTATED, 5, 0.1, 0.0, 1, 0.5, id="one-rung-top"),
    pytest.param(0, O.ROTATED, 7, 0.1, 0.0, 4, 0.0, id="rotated7-no-logicals"),
    pytest.param(0, O.XZZX, 7, 0.1, 0.0, 32, 0.5, id="xzzx7-32rungs"),
    pytest.param(2, O.XZZX, 9, 0.15, 30.0, 9, 0.5, id="xzzx9-biased"),
    pytest.param(2, O.XZZX, 21, 0.15, 100.0, 21, 0.5, id="xzzx21-biased-u64"),
    pytest.param(2, O.XZZX, 7, 0.7, 4.0, 7, 0.5, id="xzzx7-biased-falling-ladder"),
    pytest.param(2, O.PLANAR, 5, 0.1, 3.0, 4, 0.5, id="planar5-biased"),
    pytest.param(1, O.XZZX, 7, 0.17, 0.65, 7, 0.5, id="xzzx7-alpha"),
    pytest.param(1, O.ROTATED, 5, 0.2, 2.0, 5, 0.5, id="rotated5-alpha"),
]


@pytest.mark.parametrize("kind,g,L,bottom,b,Nc,p_logical", LADDERS)
def test_native_ladder_steps_equal_oracle(ctx, kind, g, L, bottom, b, Nc, p_logical):
    """Ladder.step (mcmc.py:94-103 and the alpha / biased variants) on native words: after `steps` steps every rung's
    lattice, the flags, tops0 and the rung-owned n_eff equal the oracle's, ladder by ladder."""
    S, steps, iters, seed = 70, 40 if L < 20 else 12, 10, 99 + L
    rng = np.random.default_rng(500 + 3 * kind + L)
    qm = np.stack([rand_lattice(rng, g, L, 0.1).reshape(-1) for _ in range(S)])
    out = ctx.ladder_run(g, L, kind, qm, bottom, Nc, steps, iters=iters, param_b=b, p_logical=p_logical, seed=seed)
    for s in (0, 1, 31, 32, 47, 69):
        lad = O.Ladder(kind, g, L, qm[s], bottom, Nc, p_logical, b)
        st = O.Stream.ladder_native(seed, s)
        for _ in range(steps):
            lad.step(iters, st, st)
        assert np.array_equal(out["rung_states"][s], lad.qm), f"rung states differ (ladder {s})"
        assert np.array_equal(out["flags"][s], lad.flags) and out["tops0"][s] == lad.tops0.value
        if kind == 1:
            assert np.array_equal(out["n_eff"][s], lad.n_eff)
    assert out["stats"]["accepted"] > 0


@pytest.mark.parametrize("kind,g,L,bottom,b,Nc,cap", [(0, O.ROTATED, 7, 0.12, 0.0, None, 6000), (0, O.TORIC, 5, 0.1, 0.0, None, 20000),
                                                      (0, O.ROTATED, 25, 0.12, 0.0, 25, 300), (2, O.XZZX, 21, 0.12, 100.0, 21, 300),
                                                      (2, O.XZZX, 7, 0.12, 10.0, None, 6000), (1, O.XZZX, 7, 0.15, 2.0, None, 6000),
                                                      (0, O.PLANAR, 5, 0.1, 0.0, None, 20000)])
def test_native_pteq_equals_oracle(ctx, kind, g, L, bottom, b, Nc, cap):
    """PTEQ / PTEQ_biased / PTEQ_alpha (decoders.py:25-105, decoders_biasednoise.py:28-237) on native words: steps used,
    since_burn, tops0, the class counts and the uint8 percentages equal the oracle's for every ladder checked -- including
    ladders that started when an earlier one converged (the grid is capped so that ladders queue up)."""
    S, seed = 200, 4242 + L
    rng = np.random.default_rng(800 + kind + L)
    qm = np.stack([rand_lattice(rng, g, L, 0.08).reshape(-1) for _ in range(S)])
    ctx.debug_set("pt_grid", 2)
    try:
        pct, info = ctx.pteq(g, L, kind, qm, bottom, Nc=Nc, param_b=b, steps=cap, seed=seed)
    finally:
        ctx.debug_set("pt_grid", -1)
    check = (0, 1, 33, 64, 65, 130, 199)

    def run(s):
        st = O.Stream.ladder_native(seed, s)
        return O.pteq(kind, g, L, qm[s], bottom, st, st, Nc=Nc, param_b=b, steps=cap)

    want = parallel(run, check)
    for s in check:
        w_pct, w = want[s]
        assert info["steps"][s] == w["steps"], (s, info["steps"][s], w["steps"])
        assert info["since_burn"][s] == w["since_burn"] and info["tops0"][s] == w["tops0"]
        assert np.array_equal(info["counts"][s], w["counts"])
        assert np.array_equal(pct[s], w_pct)
    if cap >= 6000:
        assert 0 < info["converged"].sum() < S      # some ladders converged before the cap (and started their successors), others ran into it


def test_native_pteq_does_not_depend_on_the_grid(ctx):
    """Ladders are seeded by their index in the call: whichever CTA picks a ladder up, and whenever, the result is the same."""
    g, L, S = O.ROTATED, 5, 300
    rng = np.random.default_rng(77)
    qm = np.stack([rand_lattice(rng, g, L, 0.08).reshape(-1) for _ in range(S)])
    a = ctx.pteq(g, L, 0, qm, 0.1, steps=20000, seed=3)
    ctx.debug_set("pt_grid", 1)
    try:
        b = ctx.pteq(g, L, 0, qm, 0.1, steps=20000, seed=3)
    finally:
        ctx.debug_set("pt_grid", -1)
    assert np.array_equal(a[0], b[0])
    for k in ("steps", "since_burn", "tops0", "counts", "converged"):
        assert np.array_equal(a[1][k], b[1][k]), k


@pytest.mark.parametrize("lt", [2, 8, 32])
def test_top_rung_lane_split_does_not_change_results(ctx, lt):
    """The number of lanes sharing a top-rung replica is a tuning knob: any value gives the same chains."""
    g, L, S = O.TORIC, 7, 40
    rng = np.random.default_rng(5)
    qm = np.stack([rand_lattice(rng, g, L, 0.1).reshape(-1) for _ in range(S)])
    a = ctx.ladder_run(g, L, 0, qm, 0.1, 7, 30, seed=8, p_logical=0.5)
    ctx.debug_set("pt_lt", lt)
    try:
        b = ctx.ladder_run(g, L, 0, qm, 0.1, 7, 30, seed=8, p_logical=0.5)
    finally:
        ctx.debug_set("pt_lt", -1)
    assert np.array_equal(a["rung_states"], b["rung_states"]) and np.array_equal(a["tops0"], b["tops0"])
