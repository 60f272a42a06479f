"""GPU parity: wt_huff<> rank / select / access / inverse_select through the C ABI,
against the oracle and the unmodified reference — the shape of test/wt_byte_test.cpp:106-203."""
import numpy as np
import pytest

import texts

pytestmark = pytest.mark.gpu


def _checkers(oracle, orc, t):
    out = [("oracle", oracle.wt_huff(t))]
    if orc.ref_available():
        out.append(("reference", orc.Ref().wt_huff(t)))
    return out


@pytest.mark.parametrize("level_sync", [None, "0", "1"], ids=["auto", "per_query", "level_sync"])
def test_wt_huff_catalogue(pkg, oracle, orc, monkeypatch, level_sync):
    """every operation under both execution strategies (one thread per query / one launch per tree depth for rank,
    operator[], inverse_select and select), forced onto every text of the catalogue, skewed Huffman shapes included"""
    if level_sync is not None:
        monkeypatch.setenv("SDSLGPU_WT_LEVEL_SYNC", level_sync)
    rng = np.random.default_rng(77)
    for name, t in texts.text_catalogue(large=True):
        n = len(t)
        arr = np.frombuffer(t, dtype=np.uint8)
        with pkg.WtHuff(t) as wt:
            assert wt.size == n and wt.sigma == len(set(t)), name
            i, c = texts.wt_queries(t, rng, min(60000, 20 * n + 16))
            j = rng.integers(0, n, len(i), dtype=np.uint64)
            got_rank = wt.rank(i, c)
            got_rnk, got_sym = wt.inverse_select(j)
            assert (wt.access(j) == arr[j.astype(np.int64)]).all(), (name, "access vs text")
            if n:  # out of domain stays NPOS through all passes
                jb = j.copy()
                jb[::7] = n + 3
                ab = wt.access(jb)
                rb, sb = wt.inverse_select(jb)
                assert (ab[::7] == pkg.NPOS).all() and (rb[::7] == pkg.NPOS).all() and (sb[::7] == pkg.NPOS).all(), (name, "access ood")
                keep = np.arange(len(j)) % 7 != 0
                assert (ab[keep] == arr[j[keep].astype(np.int64)]).all() and (rb[keep] == got_rnk[keep]).all(), (name, "access mixed")
            # rank(size, c) for all 256 symbols == histogram (test/wt_byte_test.cpp:160-167)
            allc = np.arange(256, dtype=np.uint8)
            tot = wt.rank(np.full(256, n, dtype=np.uint64), allc)
            assert (tot == np.bincount(arr, minlength=256).astype(np.uint64)).all(), (name, "rank(size,c)")
            occ = tot[c.astype(np.int64)]
            ok = occ > 0
            k = (rng.integers(0, 2**62, len(i), dtype=np.uint64) % np.maximum(occ, 1)) + np.uint64(1)
            got_sel = wt.select(k[ok], c[ok])
            for cname, chk in _checkers(oracle, orc, t):
                assert (got_rank == chk.rank(i, c)).all(), (name, cname, "rank")
                rr, ss = chk.inverse_select(j)
                assert (got_rnk == rr).all() and (got_sym == ss).all(), (name, cname, "inverse_select")
                assert (got_sel == chk.select(k[ok], c[ok])).all(), (name, cname, "select")
            # select of an absent symbol returns size(); beyond the occurrences is NPOS (defined here)
            absent = np.nonzero(tot == 0)[0].astype(np.uint8)
            if len(absent):
                assert (wt.select(np.ones(len(absent), np.uint64), absent) == n).all()
            present = np.nonzero(tot > 0)[0].astype(np.uint8)
            beyond = wt.select(tot[present.astype(np.int64)] + np.uint64(1), present)
            # (for sigma == 1 the reference's in-band answer min(i-1, size) is kept, wt_pc.hpp:451-454)
            assert (beyond == (n if wt.sigma == 1 else pkg.NPOS)).all(), (name, "select beyond")
            # round trip: select(rank(j, wt[j]) + 1, wt[j]) == j
            assert (wt.select(got_rnk + np.uint64(1), got_sym.astype(np.uint8)) == j).all(), (name, "round trip")


def test_wt_huff_large_properties(pkg):
    """2^26-symbol uniform text (the shape of BASELINE config 4, scaled): histogram and round-trip properties"""
    rng = np.random.default_rng(5)
    n = 1 << 26
    t = rng.integers(0, 256, n, dtype=np.uint8)
    with pkg.WtHuff(t) as wt:
        tot = wt.rank(np.full(256, n, dtype=np.uint64), np.arange(256, dtype=np.uint8))
        assert (tot == np.bincount(t, minlength=256).astype(np.uint64)).all()
        j = rng.integers(0, n, 1_000_000, dtype=np.uint64)
        rnk, sym = wt.inverse_select(j)
        assert (sym == t[j.astype(np.int64)]).all()
        assert (wt.select(rnk + np.uint64(1), sym.astype(np.uint8)) == j).all()
        assert (wt.rank(j, sym.astype(np.uint8)) == rnk).all()
        # rank is monotone in i and increases by exactly [t[i] == c]
        c = sym.astype(np.uint8)
        assert (wt.rank(j + np.uint64(1), c) == rnk + np.uint64(1)).all()


@pytest.mark.parametrize("mode", ["0", "1"])
def test_wt_rank_level_synchronous_equals_per_query(pkg, oracle, monkeypatch, mode):
    """both rank strategies (per-query kernel / one launch per tree depth) forced onto every text of the catalogue,
    skewed Huffman shapes included, and compared with the oracle"""
    monkeypatch.setenv("SDSLGPU_WT_LEVEL_SYNC", mode)
    rng = np.random.default_rng(78)
    for name, t in texts.text_catalogue(large=True):
        with pkg.WtHuff(t) as wt:
            i, c = texts.wt_queries(t, rng, min(80000, 20 * len(t) + 16))
            want = oracle.wt_huff(t).rank(i, c)
            i = i.copy()
            i[1] = len(t) + 5  # out of domain stays NPOS through all passes
            got = wt.rank(i, c)
            assert got[1] == pkg.NPOS
            got[1] = want[1]
            assert (got == want).all(), (name, mode)
