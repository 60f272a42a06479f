"""CPU: the C restatement of wt_huff<> and csa_wt<wt_huff<>> (oracle/oracle_wt.c, oracle_csa.c) against the
UNMODIFIED reference (oracle/_ref) — rows a7, a9, a10.  Serialised bytes pin construction (Huffman shape,
BFS layout, bit planes, rank/select supports, SA/ISA samples, alphabet); queries pin rank / select /
inverse_select / count / locate / SA access.  count() and locate() on the byte FM-index are NOT pinned by the
reference's own tests, so they are additionally checked against a brute-force scan here."""
import numpy as np
import pytest

import texts


def _patterns(t, rng, k):
    n = len(t)
    pats = []
    for _ in range(k):
        m = int(rng.integers(1, 25))
        if n and rng.random() < 0.7:
            s = int(rng.integers(0, max(1, n - m + 1)))
            pats.append(t[s : s + m])
        else:
            pats.append(rng.integers(1, 256, m, dtype=np.uint8).tobytes())
    return pats + [b"", t, t[:4], t + b"x"]


def test_wt_huff_vs_reference(oracle, ref):
    rng = np.random.default_rng(3)
    for name, t in texts.text_catalogue(large=False):
        ow, rw = oracle.wt_huff(t), ref.wt_huff(t)
        assert ow.serialize() == rw.serialize(), (name, "serialised bytes")
        n = len(t)
        i, c = texts.wt_queries(t, rng, min(20000, 8 * n + 16))
        assert (ow.rank(i, c) == rw.rank(i, c)).all(), (name, "rank")
        tot = rw.rank(np.full(len(c), n, dtype=np.uint64), c)
        ok = tot > 0
        k = (rng.integers(0, 2**62, len(c), dtype=np.uint64) % np.maximum(tot, 1)) + np.uint64(1)
        assert (ow.select(k[ok], c[ok]) == rw.select(k[ok], c[ok])).all(), (name, "select")
        j = rng.integers(0, n, len(c), dtype=np.uint64)
        a, b = ow.inverse_select(j), rw.inverse_select(j)
        assert (a[0] == b[0]).all() and (a[1] == b[1]).all(), (name, "inverse_select")


def test_wt_huff_vs_naive(oracle):
    """test/wt_byte_test.cpp:134-184: prefix counts and k-th occurrences against the raw text"""
    for name, t in texts.text_catalogue(large=False):
        arr = np.frombuffer(t, dtype=np.uint8)
        n = len(arr)
        if n > 50000:
            continue
        w = oracle.wt_huff(t)
        for c in sorted(set(t))[:6] + [255 if 255 not in t else 254]:
            hit = arr == c
            pref = np.concatenate([[0], np.cumsum(hit)]).astype(np.uint64)
            assert (w.rank(np.arange(n + 1, dtype=np.uint64), np.full(n + 1, c, np.uint8)) == pref).all(), (name, c)
            pos = np.nonzero(hit)[0].astype(np.uint64)
            if len(pos):
                assert (w.select(np.arange(1, len(pos) + 1, dtype=np.uint64), np.full(len(pos), c, np.uint8)) == pos).all()
        assert (w.access(np.arange(n, dtype=np.uint64)) == arr).all(), name


def test_csa_vs_reference(oracle, ref, orc):
    rng = np.random.default_rng(4)
    for name, t in texts.text_catalogue(zero_free=True, large=False):
        oc, rc = oracle.csa(t), ref.csa(t)
        assert oc.serialize() == rc.serialize(), (name, "serialised bytes")
        flat, off = orc.csr_patterns(_patterns(t, rng, 600))
        c1, l1 = oc.count(flat, off, want_l=True)
        c2, l2 = rc.count(flat, off, want_l=True)
        assert (c1 == c2).all() and (l1[c1 > 0] == l2[c1 > 0]).all(), (name, "count / interval")
        a1, a2 = oc.locate(flat, off), rc.locate(flat, off)
        assert (a1[0] == a2[0]).all() and (a1[1] == a2[1]).all(), (name, "locate (SA order)")
        i = rng.integers(0, len(t) + 1, 2000, dtype=np.uint64)
        assert (oc.sa(i) == rc.sa(i)).all(), (name, "SA access")


def test_count_locate_vs_bruteforce(oracle, orc):
    rng = np.random.default_rng(6)
    for name, t in texts.text_catalogue(zero_free=True, large=False):
        if len(t) > 60000:
            continue
        oc = oracle.csa(t)
        pats = _patterns(t, rng, 150)
        flat, off = orc.csr_patterns(pats)
        cnt = oc.count(flat, off)
        occ_off, occ = oc.locate(flat, off)
        for k, p in enumerate(pats):
            if len(p) == 0:
                assert cnt[k] == len(t) + 1  # empty pattern: the whole suffix-array interval
                continue
            want = []
            s = t.find(p)
            while s >= 0:
                want.append(s)
                s = t.find(p, s + 1)
            assert cnt[k] == len(want), (name, p)
            assert sorted(occ[int(occ_off[k]) : int(occ_off[k + 1])].tolist()) == want, (name, p)
