"""csrc/qecmc_lattice.h (the packed-lattice geometry used by the CUDA kernels), compiled
for the host, against the oracle: every stabilizer, every logical, classes, to_class,
the canonical index maps and the linear fingerprint."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "harness", "lattice_harness.cpp")
SO = os.path.join(HERE, "harness", "_lattice_harness.so")


@pytest.fixture(scope="module")
def lh():
    hdr = os.path.join(HERE, "..", "mcmc-qec-toric-rl_b200", "csrc", "qecmc_lattice.h")
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(SRC), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O1", "-shared", "-fPIC", "-o", SO, SRC])
    lib = C.CDLL(SO)
    lib.lh_run.argtypes = [C.c_int, C.c_int, C.c_int, np.ctypeslib.ndpointer(np.uint8, flags="C")] + [C.c_int] * 5
    return lib


def rand_lattice(rng, g, L, p=0.4):
    shape = (2, L, L) if g in (O.TORIC, O.PLANAR) else (L, L)
    return ((rng.random(shape) < p) * rng.integers(1, 4, shape)).astype(np.uint8)


def all_stabs(g, L):
    if g == O.TORIC:
        return [(r, c, op) for op in (1, 3) for r in range(L) for c in range(L)]
    if g == O.PLANAR:
        return [(r, c, 1) for r in range(L - 1) for c in range(L)] + [(r, c, 3) for r in range(L) for c in range(L - 1)]
    return [(r, c, 1) for r in range(L - 1) for c in range(L - 1)] + [(k, s, 3) for k in range((L - 1) // 2) for s in range(4)]


CONFIGS = [(g, L, wide) for g in range(4) for (L, wide) in ((3, 0), (5, 0), (15, 0), (16, 0) if g < 2 else (13, 0), (5, 1), (21, 1), (25, 1))]


@pytest.mark.parametrize("g,L,wide", CONFIGS)
def test_packed_geometry(lh, g, L, wide):
    rng = np.random.default_rng(100 * g + L + wide)
    q = rand_lattice(rng, g, L)
    stabs = all_stabs(g, L)
    assert lh.lh_nstab(g, L) == len(stabs)
    seen = set()
    for (r, c, op) in stabs:
        want, wd = O.apply_stabilizer(g, L, q, r, c, op)
        got = q.reshape(-1).copy()
        d = lh.lh_run(g, L, wide, got, 0, r, c, op, 0)
        assert d == wd and np.array_equal(got.reshape(q.shape), want), (r, c, op)
    for idx in range(len(stabs)):
        assert lh.lh_run(g, L, wide, q.reshape(-1).copy(), 6, idx, 0, 0, 0) == idx
        got = q.reshape(-1).copy()
        lh.lh_run(g, L, wide, got, 1, idx, 0, 0, 0)
        seen.add(got.tobytes())
        assert lh.lh_run(g, L, wide, q.reshape(-1).copy(), 8, idx, 0, 0, 0) == 1
    # the canonical index enumerates every stabilizer exactly once
    assert seen == {O.apply_stabilizer(g, L, q, *s)[0].tobytes() for s in stabs}
    for op in range(4):
        for layer in ((0, 1) if g == O.TORIC else (0,)):
            for xp in range(L):
                zp = (2 * xp + 1) % L
                want, wd = O.apply_logical(g, L, q, op, layer, xp, zp)
                got = q.reshape(-1).copy()
                d = lh.lh_run(g, L, wide, got, 2, op, layer, xp, zp)
                assert d == wd and np.array_equal(got.reshape(q.shape), want)
    for _ in range(10):
        x = rand_lattice(rng, g, L, 0.3)
        assert lh.lh_run(g, L, wide, x.reshape(-1).copy(), 4, 0, 0, 0, 0) == O.eq_class(g, L, x)
        assert lh.lh_run(g, L, wide, x.reshape(-1).copy(), 5, 0, 0, 0, 0) == int((x != 0).sum())
        for e in range(O.neq(g)):
            got = x.reshape(-1).copy()
            lh.lh_run(g, L, wide, got, 3, e, 0, 0, 0)
            assert np.array_equal(got.reshape(x.shape), O.to_class(g, L, x, e))
