"""CPU tests of the host side: the code objects mirror the reference's interface and agree with the oracle
(itself pinned to the reference's golden vectors), the C-ABI library loads and exports every symbol that
include/qecmc.h declares, and the decoder mirrors keep the reference's signatures.  No compute calls."""
import ctypes
import inspect
import os
import re
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import oracle as O  # noqa: E402
from mcmc_qec_toric_rl_b200 import _lib, decoders, decoders_biasednoise  # noqa: E402
from mcmc_qec_toric_rl_b200.src.toric_model import Toric_code  # noqa: E402
from mcmc_qec_toric_rl_b200.src.planar_model import Planar_code  # noqa: E402
from mcmc_qec_toric_rl_b200.src.rotated_surface_model import RotSurCode  # noqa: E402
from mcmc_qec_toric_rl_b200.src.xzzx_model import xzzx_code  # noqa: E402

CODES = [(Toric_code, O.TORIC, 5), (Toric_code, O.TORIC, 6), (Planar_code, O.PLANAR, 5), (RotSurCode, O.ROTATED, 5),
         (RotSurCode, O.ROTATED, 7), (xzzx_code, O.XZZX, 5), (xzzx_code, O.XZZX, 7)]


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "qecmc.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(qecmc_[a-z0-9_]+)\s*\(", hdr))
    assert {"qecmc_stdc", "qecmc_strc", "qecmc_pteq", "qecmc_ladder_run", "qecmc_stdc_alpha", "qecmc_ptdc",
            "qecmc_single_temp", "qecmc_replay_chain", "qecmc_chain_update"} <= names
    lib = ctypes.CDLL(_lib.SO)
    missing = [n for n in sorted(names) if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.qecmc_abi_version() == 1


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(_lib.QecmcError):
        _lib.Context(0)


@pytest.mark.parametrize("cls,g,L", CODES)
def test_code_object_matches_oracle(cls, g, L):
    rng = np.random.default_rng(10 * g + L)
    code = cls(L)
    assert code.nbr_eq_classes == O.neq(g) and code.qubit_matrix.dtype == np.uint8
    shape = code.qubit_matrix.shape
    q = ((rng.random(shape) < 0.3) * rng.integers(1, 4, shape)).astype(np.uint8)
    if g == O.PLANAR:
        q[1, -1, :] = 0
        q[1, :, -1] = 0
    code.qubit_matrix = q.copy()
    assert code.count_errors() == int((q != 0).sum())
    assert code.define_equivalence_class() == O.eq_class(g, L, q)
    # every stabilizer: same new lattice and weight change, the object itself untouched (functional style)
    for (r, c, op) in code._stabilizer_table(L):
        new, d = code.apply_stabilizer(r, c, op)
        want, wd = O.apply_stabilizer(g, L, q, r, c, op)
        assert np.array_equal(new, want) and d == wd
    assert np.array_equal(code.qubit_matrix, q)
    # logical operators and class moves
    for op in range(4):
        for pos in (0, L - 1):
            if g == O.TORIC:
                for layer in (0, 1):
                    new, d = code.apply_logical(op, layer, pos, (pos + 1) % L)
                    want, wd = O.apply_logical(g, L, q, op, layer, pos, (pos + 1) % L)
                    assert np.array_equal(new, want) and d == wd
            else:
                new, d = code.apply_logical(op, pos, (pos + 1) % L)
                want, wd = O.apply_logical(g, L, q, op, 0, pos, (pos + 1) % L)
                assert np.array_equal(new, want) and d == wd
    for eq in range(code.nbr_eq_classes):
        moved = code.to_class(eq)
        assert np.array_equal(moved, O.to_class(g, L, q, eq))
    # stabilizers commute with the syndrome: applying one never changes it
    syn = code.syndrome()
    r, c, op = next(iter(code._stabilizer_table(L)))
    code.qubit_matrix = code.apply_stabilizer(r, c, op)[0]
    assert code.syndrome() == syn


@pytest.mark.parametrize("cls,g,L", CODES)
def test_random_moves_keep_syndrome_and_report_delta(cls, g, L):
    np.random.seed(3)
    code = cls(L)
    code.generate_random_error(0.2) if g != O.TORIC else code.generate_random_error(0.2)
    syn = code.syndrome()
    for _ in range(20):
        before = code.count_errors()
        new, d = code.apply_random_stabilizer()
        code.qubit_matrix = new
        assert code.count_errors() - before == d
    assert code.syndrome() == syn
    new, d = code.apply_random_logical()
    assert int((new != 0).sum()) - code.count_errors() == d


def test_rain_only_on_two_layer_codes():
    np.random.seed(0)
    for cls, g in ((Toric_code, O.TORIC), (Planar_code, O.PLANAR)):
        code = cls(5)
        code.generate_random_error(0.1)
        syn = code.syndrome()
        code.qubit_matrix = code.apply_stabilizers_uniform()
        assert code.syndrome() == syn
    with pytest.raises(AttributeError):
        RotSurCode(5).apply_stabilizers_uniform()


REFERENCE_SIGNATURES = {
    # name: positional parameter names of the reference (decoders.py / decoders_biasednoise.py)
    "STDC": ["init_code", "p_error", "p_sampling", "droplets", "steps", "conv_mult"],
    "STRC": ["init_code", "p_error", "p_sampling", "droplets", "steps", "conv_mult"],
    "PTEQ": ["init_code", "p", "Nc", "SEQ", "TOPS", "tops_burn", "eps", "steps", "iters", "conv_criteria"],
    "PTDC": ["init_code", "p_error", "p_sampling", "droplets", "Nc", "steps", "conv_mult"],
    "single_temp": ["init_code", "p", "max_iters"],
    "STDC_Nall_n_alpha": ["init_code", "pz_tilde_sampling", "alpha", "pz_tilde", "steps"],
}
DEFAULTS = {"STDC": (None, 10, 20000, 0), "STRC": (None, 10, 20000, 0), "PTDC": (None, 4, None, 20000, 0),
            "PTEQ": (None, 2, 10, 2, 0.1, 50000000, 10, 'error_based'), "STDC_Nall_n_alpha": (None, 1, 0.1, 20000)}


@pytest.mark.parametrize("name", sorted(REFERENCE_SIGNATURES))
def test_decoder_signatures_match_reference(name):
    sig = inspect.signature(getattr(decoders, name))
    assert list(sig.parameters) == REFERENCE_SIGNATURES[name]
    if name in DEFAULTS:
        got = tuple(p.default for p in sig.parameters.values() if p.default is not inspect.Parameter.empty)
        assert got == DEFAULTS[name]


def test_biased_decoder_signatures_match_reference():
    sig = inspect.signature(decoders_biasednoise.PTEQ_biased)
    assert list(sig.parameters) == ["init_code", "p", "eta", "Nc", "SEQ", "TOPS", "tops_burn", "eps", "steps", "iters",
                                    "conv_criteria"]
    assert sig.parameters["eta"].default == 0.5
    sig = inspect.signature(decoders_biasednoise.PTEQ_alpha)
    assert list(sig.parameters)[:4] == ["init_code", "pz_tilde", "alpha", "Nc"] and sig.parameters["alpha"].default == 1


def test_dtype_is_enforced_like_the_reference():
    code = Toric_code(5)
    code.qubit_matrix = np.zeros((2, 5, 5), np.int64)   # the njit signatures reject anything but uint8
    with pytest.raises(TypeError):
        decoders.STDC(code, 0.1, steps=10)
    with pytest.raises(TypeError):
        decoders.STDC(RotSurCode(5), 0.1, steps=10)     # 2-D lattice into the 3-D-only fast path


def test_generate_data_noise_models_match_the_reference_formulas():
    """generate_data.py:57-118: per-qubit Pauli probabilities of the four noise models"""
    from mcmc_qec_toric_rl_b200 import generate_data as G
    assert G.noise_probabilities({'code': 'toric', 'noise': 'depolarizing', 'p_error': 0.12}) == (0.12, None)
    pe, (px, py, pz) = G.noise_probabilities({'code': 'planar', 'noise': 'depolarizing', 'p_error': 0.12})
    assert pe is None and px == py == pz == 0.12 / 3
    _, (px, py, pz) = G.noise_probabilities({'code': 'xzzx', 'noise': 'biased', 'p_error': 0.15, 'eta': 100})
    assert np.isclose(pz, 0.15 * 100 / 101) and np.isclose(px, 0.15 / 202) and px == py
    _, (px, py, pz) = G.noise_probabilities({'code': 'rotated', 'noise': 'alpha', 'p_error': 0.1, 'alpha': 2})
    p_tilde = 0.1 + 2 * 0.1 ** 2
    p = p_tilde / (1 + p_tilde)
    assert np.isclose(pz, 0.1 * (1 - p)) and np.isclose(px, 0.01 * (1 - p)) and px == py
    with pytest.raises(AssertionError):
        G.noise_probabilities({'code': 'toric', 'noise': 'biased', 'p_error': 0.1})


# ------------------------------------------------------------------ host-side mirrors that need no device
def _ladder_stub(cls, *args):
    """A ladder object with its rung chains but without touching the device (construction is host-only)."""
    return cls(*args)


def test_ladder_r_flip_follows_the_reference_rules(monkeypatch):
    """Ladder.r_flip (src/mcmc.py:86-92, 144-149): a lighter upper replica swaps without a draw, otherwise
    u < p_diff ** (ne_hi - ne_lo); Ladder_biased.r_flip (mcmc_biased.py:107-113) and Ladder_alpha.r_flip
    (mcmc_alpha.py:117-123) always draw."""
    import random
    from mcmc_qec_toric_rl_b200.src.mcmc import Ladder
    from mcmc_qec_toric_rl_b200.src.mcmc_biased import Ladder_biased
    from mcmc_qec_toric_rl_b200.src.mcmc_alpha import Ladder_alpha
    code = Planar_code(5)
    lad = Ladder(0.1, code, 3, 0.5)
    lad.chains[0].code.qubit_matrix[0, 0, :3] = 1      # 3 errors on rung 0
    lad.chains[1].code.qubit_matrix[0, 1, :1] = 2      # 1 error on rung 1
    draws = []
    monkeypatch.setattr(random, "random", lambda: draws.append(1) or 0.999999)
    assert lad.r_flip(0) is True and not draws                       # ne_hi < ne_lo: no draw
    lad.chains[1].code.qubit_matrix[0, 1, :] = 2                     # 5 errors: now a draw decides
    want = 0.999999 < lad.p_diff[0] ** (5 - 3)
    assert lad.r_flip(0) == want and len(draws) == 1
    monkeypatch.setattr(random, "random", lambda: 0.0)
    assert lad.r_flip(0) is True
    lb = Ladder_biased(0.1, xzzx_code(5), 10.0, 3, 0.5)
    lb.chains[0].code.qubit_matrix[0, :3] = 3
    monkeypatch.setattr(random, "random", lambda: draws.append(1) or 0.5)
    n0 = len(draws)
    assert lb.r_flip(0) == (0.5 < lb.p_diff[0] ** (0 - 3)) and len(draws) == n0 + 1   # draws even though ne_hi < ne_lo
    la = Ladder_alpha(0.1, xzzx_code(5), 2.0, 3, 0.5)
    la.chains[1].n_eff = 4.0
    assert la.r_flip(0) == (0.5 < (la.chains[0].pz_tilde / la.chains[1].pz_tilde) ** 4.0)


def test_chain_mirrors_keep_the_reference_attributes():
    from mcmc_qec_toric_rl_b200.src.mcmc_biased import Chain_biased
    from mcmc_qec_toric_rl_b200.src.mcmc_alpha import Chain_alpha
    cb = Chain_biased(0.12, 100.0, xzzx_code(5))
    assert cb.factor == (0.12 / 3.0) / (1.0 - 0.12) and callable(cb.update_chain_fast)      # mcmc_biased.py:17, 62-63
    with pytest.raises(AttributeError):                                                     # mcmc_alpha.py:21, 73-74
        Chain_alpha(0.1, 2.0, xzzx_code(5)).update_chain_fast(10)


def test_call_keys_do_not_collide_or_overflow():
    """The Philox key of call k of a host-driven ladder: distinct across calls and seeds, always < 2^64
    (the old stream + (calls << 44) overlapped the seed bits and wrapped after 2^20 calls)."""
    from mcmc_qec_toric_rl_b200.src import mcmc
    keys = {mcmc._call_key(s, k) for s in (1 << 24, (1 << 24) + (1 << 44), 12345) for k in (0, 1, 2, 1 << 20, (1 << 20) + 1, 5 * 10**7)}
    assert len(keys) == 18 and all(0 <= k < 1 << 64 for k in keys)


def test_generate_alpha_error_rates():
    """planar_model.py:79-99: pz_tilde solves x + 2 x^alpha = p/(1+p); Z dominates for alpha > 1."""
    np.random.seed(3)
    for code in (Planar_code(15), xzzx_code(15)):
        code.generate_alpha_error(0.3, 2.0)
        nx, ny, nz = code.chain_lengths()
        assert code.qubit_matrix.dtype == np.uint8 and nz > 2 * (nx + ny) and nz > 0
    pc = Planar_code(7)
    pc.generate_alpha_error(0.5, 1.0)
    assert not pc.qubit_matrix[1, -1, :].any() and not pc.qubit_matrix[1, :, -1].any()


def test_data_reader_raises_on_a_missing_file(tmp_path):
    from mcmc_qec_toric_rl_b200.src.mcmc import MCMCDataReader
    with pytest.raises(FileNotFoundError):
        MCMCDataReader(str(tmp_path / "nope.xz"), 5)


def test_convergence_criterion_helper():
    """conv_crit_error_based_PT (decoders.py:91-104) on host arrays: quarters 2 and 4 of the valid part of the history."""
    from mcmc_qec_toric_rl_b200 import decoders, decoders_biasednoise
    hist = np.zeros(64)
    hist[:40] = [5] * 10 + [7] * 10 + [9] * 10 + [7] * 10          # since_burn = 39: Q2 = hist[10:20] = 7, Q4 = hist[30:40] = 7
    assert decoders.conv_crit_error_based_PT(hist, 39, 3, 2, 0.1) == (True, True)
    assert decoders.conv_crit_error_based_PT(hist, 39, 1, 2, 0.1) == (True, False)
    hist[30:40] = 7.2
    assert decoders.conv_crit_error_based_PT(hist, 39, 3, 2, 0.1) == (False, False)
    assert decoders.conv_crit_error_based_PT(hist, 39, 3, 2, 0.25) == (True, True)
    assert decoders.conv_crit_error_based_PT(hist, 0, 3, 2, 0.1) == (False, False)     # one entry: empty quarters
    assert decoders_biasednoise.conv_crit_error_based_PT_alpha is decoders.conv_crit_error_based_PT
    assert decoders_biasednoise.conv_crit_error_based_PT_biased is decoders.conv_crit_error_based_PT
