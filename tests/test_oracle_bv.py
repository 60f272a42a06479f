"""CPU: the C restatement (oracle/oracle.c) against the UNMODIFIED reference (oracle/_ref) — rows a1..a4.

Byte-exact serialisation parity pins the builders; query parity pins rank/select.
"""
import numpy as np
import pytest

import cases


def test_bits(oracle):
    L = oracle.L
    rng = np.random.default_rng(3)
    for i in range(64):
        assert L.orc_cnt(1 << i) == 1 and L.orc_hi(1 << i) == i and L.orc_lo(1 << i) == i
    assert L.orc_hi(0) == 0 and L.orc_lo(0) == 0  # bits.hpp:653,689 conventions
    for x in rng.integers(1, 2**64, 300, dtype=np.uint64):
        x = int(x)
        pos = [i for i in range(64) if (x >> i) & 1]
        assert L.orc_cnt(x) == len(pos)
        for k, p in enumerate(pos, 1):  # test/bits_test.cpp:150-177: sel vs naive walk
            assert L.orc_sel(x, k) == p


@pytest.mark.parametrize("large", [False, True])
def test_rank_select_vs_reference(oracle, ref, large):
    total = 0
    for cid, w, nbits in cases.bitvector_catalogue(large=large):
        if large and nbits < 100000:
            continue
        ob, rb = oracle.bv(w, nbits), ref.bv(w, nbits)
        for what in range(5):  # bit_vector, rank_v<1>, rank_v<0>, select_mcl<1>, select_mcl<0>
            assert ob.serialize(what) == rb.serialize(what), (cid, what)
        idx = cases.rank_queries(nbits, 1, 30000)
        for b in (0, 1):
            r = rb.rank(idx, b)
            assert (ob.rank(idx, b) == r).all(), (cid, "rank", b)
            m = int(rb.rank([nbits], b)[0])
            q = cases.select_queries(m, 2, 30000)
            if len(q):
                assert (ob.select(q, b) == rb.select(q, b)).all(), (cid, "select", b)
            total += len(idx) + len(q)
    assert total > 1000


def test_rank_select_vs_naive(oracle):
    """the reference's own property (test/rank_support_test.cpp:109-126, select_support_test.cpp:85-104):
    rank(j) == running count for every j; select(k) == position of the k-th occurrence"""
    for cid, w, nbits in cases.bitvector_catalogue(large=False):
        bits = cases.unpack_bits(w, nbits).astype(np.int64)
        ob = oracle.bv(w, nbits)
        for b in (0, 1):
            hit = bits == b
            pref = np.concatenate([[0], np.cumsum(hit)]).astype(np.uint64)
            assert (ob.rank(np.arange(nbits + 1, dtype=np.uint64), b) == pref).all(), cid
            pos = np.nonzero(hit)[0].astype(np.uint64)
            if len(pos):
                assert (ob.select(np.arange(1, len(pos) + 1, dtype=np.uint64), b) == pos).all(), cid


# two-bit patterns: code -> (previous bit, current bit); the position of an occurrence is that of its SECOND bit
PATTERNS = {2: (1, 0), 3: (0, 1), 4: (0, 0), 5: (1, 1)}


def pattern_hits(bits, code):
    prev, cur = PATTERNS[code]
    hit = np.zeros(len(bits), dtype=bool)
    if len(bits) > 1:
        hit[1:] = (bits[:-1] == prev) & (bits[1:] == cur)
    return hit


def test_two_bit_patterns_vs_naive_and_reference(oracle, ref):
    """rank_support_v<10|01|00|11, 2> and select_support_mcl<..., 2> (rank_support.hpp:161-284,
    select_support.hpp:204-405): the restatement against the definition the reference's tests use
    (rank_support_test.cpp:109-126, select_support_test.cpp:85-104) and against the reference itself"""
    for cid, w, nbits in cases.bitvector_catalogue(large=False):
        bits = cases.unpack_bits(w, nbits).astype(np.int64)
        ob, rb = oracle.bv(w, nbits), ref.bv(w, nbits)
        every = np.arange(nbits + 1, dtype=np.uint64) if nbits <= 20000 else cases.rank_queries(nbits, 5, 20000)
        for code in PATTERNS:
            hit = pattern_hits(bits, code)
            pref = np.concatenate([[0], np.cumsum(hit)]).astype(np.uint64)
            got = ob.rank(every, code)
            assert (got == pref[every.astype(np.int64)]).all(), (cid, code, "rank vs naive")
            assert (got == rb.rank(every, code)).all(), (cid, code, "rank vs reference")
            pos = np.nonzero(hit)[0].astype(np.uint64)
            if len(pos):
                k = np.arange(1, len(pos) + 1, dtype=np.uint64) if len(pos) <= 20000 else cases.select_queries(len(pos), 6, 20000)
                got = ob.select(k, code)
                assert (got == pos[k.astype(np.int64) - 1]).all(), (cid, code, "select vs naive")
                assert (got == rb.select(k, code)).all(), (cid, code, "select vs reference")


def test_rank_support_v5_vs_reference(oracle, ref):
    """rank_support_v5<b> (rank_support_v5.hpp:66-158; the table inside the reference's count-benchmark index,
    benchmark/indexing_count/index.config:8): serialised table byte-exact vs the unmodified reference, and its rank
    equal to rank_support_v's on every position of the small vectors / a sample of the large ones"""
    checked = 0
    for cid, w, nbits in cases.bitvector_catalogue(large=True):
        if nbits > 2_000_000:
            continue
        ob, rb = oracle.bv(w, nbits), ref.bv(w, nbits, with_select=False)
        for what in (5, 6):
            assert ob.serialize(what) == rb.serialize(what), (cid, what)
        idx = cases.rank_queries(nbits, 3, 3000 if nbits > 3000 else nbits + 1)
        for b in (0, 1):
            assert (ob.rank_v5(idx, b) == rb.rank(idx, b)).all(), (cid, "rank_v5", b)
        checked += 1
    assert checked > 30
