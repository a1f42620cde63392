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
