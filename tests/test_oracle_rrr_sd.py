"""CPU: the C restatement of rrr_vector<63> and sd_vector<> (oracle/oracle_rrr_sd.c) against the UNMODIFIED
reference (oracle/_ref) — rows a5, a6: serialised bytes (construction), rank / select / access for both bit
patterns, over the bit-vector catalogue, a density sweep, and sizes around the 63-bit block / 32-block
superblock boundaries (where the reference adds its dummy block and skips the invert bit)."""
import numpy as np
import pytest

import cases


def _check(oracle, ref, kind, cid, w, nbits):
    ob, rb = getattr(oracle, kind)(w, nbits), getattr(ref, kind)(w, nbits)
    assert ob.serialize() == rb.serialize(), (kind, cid, "serialised bytes")
    idx = cases.rank_queries(nbits, 1, 20000)
    for b in (1, 0):
        assert (ob.rank(idx, b) == rb.rank(idx, b)).all(), (kind, cid, "rank", b)
        m = int(rb.rank([nbits], b)[0])
        q = cases.select_queries(m, 2, 3000 if (kind == "sd" and b == 0) else 20000)
        if len(q):
            assert (ob.select(q, b) == rb.select(q, b)).all(), (kind, cid, "select", b)
    pos = idx[idx < nbits]
    if len(pos):
        assert (ob.access(pos) == rb.access(pos)).all(), (kind, cid, "access")


@pytest.mark.parametrize("kind", ["rrr", "sd"])
def test_compressed_vs_reference(oracle, ref, kind):
    for cid, w, nbits in cases.bitvector_catalogue(large=True):
        if nbits > 2_000_000 or (kind == "sd" and nbits == 0):
            continue
        _check(oracle, ref, kind, cid, w, nbits)
    for d in (0.01, 0.05, 0.25, 0.5, 0.9):
        nbits = (1 << 20) + 63 * 32 * 5
        _check(oracle, ref, kind, f"density {d}", cases.bernoulli_words(nbits, d, int(d * 100)), nbits)
    for nbits in (63, 63 * 32, 63 * 32 * 3, 63 * 31, 63 * 33, 2016 * 2 + 1):
        _check(oracle, ref, kind, f"edge {nbits}", cases.random_words(nbits, nbits), nbits)
        _check(oracle, ref, kind, f"edge90 {nbits}", cases.bernoulli_words(nbits, 0.9, nbits), nbits)


@pytest.mark.parametrize("kind", ["rrr", "sd"])
def test_compressed_vs_naive(oracle, kind):
    """test/rank_support_test.cpp:109-126 / select_support_test.cpp:85-104 on the compressed vectors"""
    for cid, w, nbits in cases.bitvector_catalogue(large=False):
        if nbits == 0 and kind == "sd":
            continue
        bits = cases.unpack_bits(w, nbits).astype(np.int64)
        ob = getattr(oracle, kind)(w, nbits)
        for b in (0, 1):
            hit = bits == b
            pref = np.concatenate([[0], np.cumsum(hit)]).astype(np.uint64)
            assert (ob.rank(np.arange(nbits + 1, dtype=np.uint64), b) == pref).all(), (kind, cid)
            pos = np.nonzero(hit)[0].astype(np.uint64)
            if len(pos):
                assert (ob.select(np.arange(1, len(pos) + 1, dtype=np.uint64), b) == pos).all(), (kind, cid)
        if nbits:
            assert (ob.access(np.arange(nbits, dtype=np.uint64)) == bits).all(), (kind, cid)
