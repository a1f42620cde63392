"""GPU tests of the reference-facing Python entry points (decoders.py, decoders_biasednoise.py, src/mcmc*.py):
return shapes / dtypes of the reference, decoding success on syndromes whose class is known, and agreement of the
decoders with each other.  These call through ctypes into libqecmc.so like a user of the reference would."""
import copy

import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    from mcmc_qec_toric_rl_b200 import decoders, decoders_biasednoise
    from mcmc_qec_toric_rl_b200.src import mcmc
    mcmc.seed(1234)
    return decoders, decoders_biasednoise, mcmc


def _workload(cls, L, p, n, seed, biased_eta=None):
    """generate_data.py:57-131: random error, remember its class, hide it behind a random logical operator."""
    np.random.seed(seed)
    import random
    random.seed(seed)
    codes, truth = [], []
    for _ in range(n):
        c = cls(L)
        if biased_eta is not None:
            c.generate_zbiased_error(p, biased_eta)
        else:
            c.generate_random_error(p)
        truth.append(c.define_equivalence_class())
        c.qubit_matrix, _ = c.apply_random_logical()
        codes.append(c)
    return codes, np.array(truth)


def test_stdc_strc_toric_d5_decode(api):
    """BASELINE config 0: toric d=5, p=0.10, STDC / STRC.  SURVEY.md section 6 measured 34/40 for the
    toric-geometry reference at steps=3000, droplets=1; the native path must be at least that good."""
    dec, _, _ = api
    codes, truth = _workload(dec.Toric_code, 5, 0.10, 100, 1)
    out = dec.STDC_batch(codes, 0.10, p_sampling=0.25, droplets=10, steps=3125, seed=5)
    assert out.shape == (100, 16) and out.dtype == np.float64
    assert np.allclose(out.sum(1), 100.0)
    ok_stdc = (out.argmax(1) == truth).mean()
    out2 = dec.STRC_batch(codes, 0.10, p_sampling=0.25, droplets=10, steps=3125, seed=6)
    ok_strc = (out2.argmax(1) == truth).mean()
    assert ok_stdc >= 0.85 and ok_strc >= 0.85, (ok_stdc, ok_strc)
    assert (out.argmax(1) == out2.argmax(1)).mean() >= 0.9
    # the single-syndrome entry point returns one row of the same thing
    one = dec.STDC(codes[0], 0.10, p_sampling=0.25, droplets=10, steps=3125)
    assert one.shape == (16,) and abs(one.sum() - 100) < 1e-9 and one.argmax() == out[0].argmax()


def test_stdc_matches_oracle_statistics_on_same_syndromes(api):
    """Native Philox STDC vs the oracle's MT19937 STDC on the same syndromes: logical failure counts within the
    binomial confidence interval of each other (north_star correctness (2))."""
    dec, _, _ = api
    n = 150
    codes, truth = _workload(dec.Toric_code, 5, 0.12, n, 2)
    gpu = dec.STDC_batch(codes, 0.12, p_sampling=0.25, droplets=4, steps=2000, seed=9)
    qm = np.stack([c.qubit_matrix.reshape(-1) for c in codes])
    ref = O.stdc_batch(O.TORIC, O.TORIC, 5, qm, 0.12, 0.25, 4, 2000, seed=17, threads=8)
    f_gpu, f_ref = (gpu.argmax(1) != truth).sum(), (ref.argmax(1) != truth).sum()
    p_hat = (f_gpu + f_ref) / (2 * n)
    sigma = np.sqrt(2 * p_hat * (1 - p_hat) / n) * n       # std of the difference of two binomial counts
    assert abs(int(f_gpu) - int(f_ref)) <= 3 * sigma + 2, (f_gpu, f_ref, sigma)
    assert (gpu.argmax(1) == ref.argmax(1)).mean() >= 0.9


def test_planar_per_class_inits_and_single_code(api):
    dec, _, _ = api
    codes, truth = _workload(dec.Planar_code, 5, 0.08, 40, 3)
    lists = []
    for c in codes:
        per_class = []
        for eq in range(4):
            k = copy.deepcopy(c)
            k.qubit_matrix = k.to_class(eq)
            per_class.append(k)
        lists.append(per_class)
    out = dec.STDC_batch(lists, 0.08, p_sampling=0.25, droplets=8, steps=2000, seed=3)
    assert out.shape == (40, 4)
    assert (out.argmax(1) == truth).mean() >= 0.85
    out_single = dec.STDC_batch(codes, 0.08, p_sampling=0.25, droplets=8, steps=2000, seed=4)
    assert (out_single.argmax(1) == out.argmax(1)).mean() >= 0.9


def test_single_temp_prefers_true_class(api):
    dec, _, _ = api
    codes, truth = _workload(dec.Toric_code, 5, 0.08, 40, 4)
    out = dec.single_temp_batch(codes, 0.08, 2000, seed=2)
    assert out.shape == (40, 16) and out.dtype == np.float64
    assert (out.argmin(1) == truth).mean() >= 0.65      # generate_data.py:199-201 takes the argmin (a weak decoder)


@pytest.mark.parametrize("cls_name,L", [("Toric_code", 5), ("Planar_code", 5), ("RotSurCode", 5), ("xzzx_code", 5)])
def test_pteq_all_codes(api, cls_name, L):
    dec, _, _ = api
    cls = getattr(dec, cls_name)
    # 96 syndromes: at p = 0.08, d = 5 the oracle on the reference's MT19937 streams decodes 85-90 % of them (toric, the
    # hardest of the four; 45 failures of 400), so 0.75 is > 3 sigma below the expected success rate
    codes, truth = _workload(cls, L, 0.08, 96, 5)
    pct, info = dec.PTEQ_batch(codes, 0.08, steps=60000, seed=11, return_info=True)
    assert pct.shape == (96, cls(L).nbr_eq_classes) and pct.dtype == np.uint8
    assert (pct.sum(1) <= 100).all() and (pct.sum(1) >= 100 - pct.shape[1]).all()      # truncation loses < 1 per class
    assert (pct.argmax(1) == truth).mean() >= 0.75
    assert info["converged"].mean() >= 0.7
    one = dec.PTEQ(codes[0], 0.08, steps=60000)
    assert one.shape == (pct.shape[1],) and one.dtype == np.uint8


def test_pteq_biased_and_alpha_xzzx(api):
    """BASELINE config 3 in small: XZZX, Z-biased noise, biased and alpha ladders."""
    dec, decb, _ = api
    eta, p = 10.0, 0.10
    codes, truth = _workload(dec.xzzx_code, 5, p, 24, 6, biased_eta=eta)
    pct = decb.PTEQ_biased_batch(codes, p, eta=eta, steps=60000, seed=1)
    assert pct.dtype == np.uint8 and pct.shape == (24, 4)
    ok_b = (pct.argmax(1) == truth).mean()
    pz_tilde = (p / (1 + 1 / eta)) / (1 - p)                 # generate_data.py:147-148
    alpha = np.log(pz_tilde / (2 * eta)) / np.log(pz_tilde)
    pct_a = decb.PTEQ_alpha_batch(codes, pz_tilde, alpha=alpha, steps=60000, seed=2)
    ok_a = (pct_a.argmax(1) == truth).mean()
    assert ok_b >= 0.75 and ok_a >= 0.75, (ok_b, ok_a)
    assert (pct.argmax(1) == pct_a.argmax(1)).mean() >= 0.8


def test_ewd_alpha_decoder(api):
    dec, _, _ = api
    eta, p = 10.0, 0.08
    codes, truth = _workload(dec.xzzx_code, 5, p, 24, 7, biased_eta=eta)
    pz_tilde = (p / (1 + 1 / eta)) / (1 - p)
    alpha = np.log(pz_tilde / (2 * eta)) / np.log(pz_tilde)
    out = dec.STDC_Nall_n_alpha_batch(codes, pz_tilde_sampling=0.3, alpha=alpha, pz_tilde=pz_tilde, steps=8000, seed=3)
    assert out.shape == (24, 4) and np.allclose(out.sum(1), 100.0)
    assert (out.argmax(1) == truth).mean() >= 0.75


def test_ptdc_agrees_with_stdc(api):
    dec, _, _ = api
    codes, truth = _workload(dec.Toric_code, 5, 0.08, 24, 8)
    pt = dec.PTDC_batch(codes, 0.08, p_sampling=0.25, droplets=2, steps=20000, seed=4)
    assert pt.dtype == np.uint8 and pt.shape == (24, 16)
    st = dec.STDC_batch(codes, 0.08, p_sampling=0.25, droplets=4, steps=4000, seed=5)
    assert (pt.argmax(1) == truth).mean() >= 0.8
    assert (pt.argmax(1) == st.argmax(1)).mean() >= 0.85


def test_chain_and_ladder_objects(api):
    """src/mcmc.py objects: same attributes as the reference's, state kept between calls, syndrome conserved."""
    dec, _, mcmc = api
    np.random.seed(9)
    code = dec.RotSurCode(5)
    code.generate_random_error(0.15)
    syn = code.syndrome()
    ch = dec.Chain(0.1, copy.deepcopy(code))
    ch.update_chain(200)
    assert ch.code.syndrome() == syn and ch.code.qubit_matrix.dtype == np.uint8
    lad = dec.Ladder(0.1, code, 5, 0.5)
    assert len(lad.chains) == 5 and np.isclose(lad.p_ladder[-1], 0.75) and lad.chains[-1].flag == 1
    for _ in range(600):
        lad.step(10)
    assert lad.tops0 > 0            # states fall from the top rung to the bottom across separate .step() calls
    assert all(c.code.syndrome() == syn for c in lad.chains)
    tor = dec.Toric_code(5)
    tor.generate_random_error(0.1)
    ch = dec.Chain(0.2, copy.deepcopy(tor))
    ch.update_chain_fast(500)
    assert ch.code.syndrome() == tor.syndrome()
    assert ch.code.define_equivalence_class() == tor.define_equivalence_class()
    la = dec.Ladder_alpha(0.2, code, 1.5, 4, 0.5)
    la.step(10)
    lb = dec.Ladder_biased(0.1, code, 3.0, 4, 0.5)
    lb.step(10)
    assert all(c.code.syndrome() == syn for c in la.chains + lb.chains)
    assert isinstance(la.chains[0].n_eff, float)
