"""CPU tests of the oracle's NATIVE mode (oracle/qec_oracle.c, "Native draws"): the Philox4x32-10 restatement against
the published known-answer vectors and an independent numpy restatement, the canonical stabilizer numbering against
the reference-pinned proposal code, and that a chain driven by native words takes the reference's own decisions."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import oracle as O  # noqa: E402

M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85


def philox_np(c0, c1, c2, c3, k0, k1):
    """Philox4x32-10 (Salmon et al., SC'11) on numpy uint64 arrays, written from the paper's round function:
    (c0, c1, c2, c3) <- (hi(M1 c2) ^ c1 ^ k0, lo(M1 c2), hi(M0 c0) ^ c3 ^ k1, lo(M0 c0)); keys bumped by the Weyl constants."""
    c0, c1, c2, c3 = [np.asarray(x, np.uint64) & np.uint64(0xFFFFFFFF) for x in (c0, c1, c2, c3)]
    k0, k1 = np.uint64(k0), np.uint64(k1)
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = c0 * np.uint64(M0), c2 * np.uint64(M1)
        c0, c1, c2, c3 = (p1 >> np.uint64(32)) ^ c1 ^ k0, p1 & mask, (p0 >> np.uint64(32)) ^ c3 ^ k1, p0 & mask
        k0, k1 = (k0 + np.uint64(W0)) & mask, (k1 + np.uint64(W1)) & mask
    return np.stack([c0, c1, c2, c3], -1).astype(np.uint32)


# Random123 kat_vectors, philox4x32 with 10 rounds
KAT = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
       ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
       ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]


@pytest.mark.parametrize("ctr,key,want", KAT)
def test_philox_known_answers(ctr, key, want):
    assert [int(x) for x in O.philox4x32_10(ctr, key)] == list(want)
    assert [int(x) for x in philox_np(*ctr, *key)] == list(want)


def test_philox_stream_order_and_numpy_restatement():
    """A native stream is the words x, y, z, w of call0, call0 + 1, ... under counter (call, tag, id lo, id hi)."""
    key, sid, tag = 0x0123456789ABCDEF, (7 << 32) | 12345, 0x80000000
    s = O.Stream.philox(key, sid, tag, call0=5)
    got = np.array([s.next() for _ in range(40)])
    calls = np.arange(5, 15)
    want = philox_np(calls, tag, sid & 0xFFFFFFFF, sid >> 32, key & 0xFFFFFFFF, key >> 32).reshape(-1)
    assert np.array_equal(got, want.astype(np.float64) / 2.0**32)
    # bits: least significant first, one word per 32 draws
    b = O.Stream.philox(key, sid, tag)
    bits = [b.next_bit() for _ in range(96)]
    w = philox_np(np.arange(1), tag, sid & 0xFFFFFFFF, sid >> 32, key & 0xFFFFFFFF, key >> 32).reshape(-1)
    assert bits == [(int(w[i >> 5]) >> (i & 31)) & 1 for i in range(96)]


@pytest.mark.parametrize("g,L", [(O.TORIC, 5), (O.TORIC, 15), (O.PLANAR, 5), (O.PLANAR, 11), (O.ROTATED, 7), (O.XZZX, 9)])
def test_native_numbering_is_a_bijection_onto_the_reference_proposals(g, L):
    """Every index names a different stabilizer, and the set equals the support of the reference's own proposal
    (toric_model.py:287-296 and friends, restated in qo_draw_stabilizer and pinned to the golden vectors)."""
    n = O.nstab(g, L)
    named = {O.stabilizer_by_index(g, L, i) for i in range(n)}
    assert len(named) == n
    mt = O.Stream.mt(5)
    seen = {O.draw_stabilizer(g, L, mt) for _ in range(60 * n)}
    assert seen == named
    # a native word w proposes index floor(w * n / 2^32)
    s = O.Stream.philox(11, 3)
    t = O.Stream.philox(11, 3)
    for _ in range(200):
        w = int(round(t.next() * 2**32))
        assert O.draw_stabilizer(g, L, s) == O.stabilizer_by_index(g, L, (w * n) >> 32)


def test_native_chain_takes_the_reference_decisions():
    """update_chain_fast under native words: the proposal comes from word 2t, and the accept rule is the reference's
    u < factor ** dE (mcmc.py:158) with u = word 2t+1 / 2^32."""
    g, L, iters, factor = O.TORIC, 5, 400, (0.25 / 3) / 0.75
    rng = np.random.default_rng(2)
    q0 = ((rng.random((2, L, L)) < 0.1) * rng.integers(1, 4, (2, L, L))).astype(np.uint8)
    out, dE, acc = O.update_chain_fast(g, L, q0, factor, iters, O.Stream.philox(9, 77), trace=True)
    words = O.Stream.philox(9, 77)
    cur = q0.copy()
    for t in range(iters):
        idx = int(words.next() * O.nstab(g, L))
        u = words.next()
        new, d = O.apply_stabilizer(g, L, cur, *O.stabilizer_by_index(g, L, idx))
        assert d == dE[t]
        a = u < O.numba_pow(factor, d)
        assert a == bool(acc[t])
        if a:
            cur = new
    assert np.array_equal(cur.reshape(-1), np.asarray(out).reshape(-1)) and 0 < acc.sum() < iters
