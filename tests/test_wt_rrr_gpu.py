"""GPU parity: wt_huff<rrr_vector<63>> and csa_wt<wt_huff<rrr_vector<63>>> (the reference's
FM_HUFF_RRR63 benchmark index, benchmark/indexing_count/index.config:10).  The tree's bit vector is stored
H0-compressed (SDSLGPU_F_RRR_BV); every answer must equal the plain-bit-vector tree's, i.e. the oracle's."""
import numpy as np
import pytest

import texts

pytestmark = pytest.mark.gpu


def test_wt_huff_over_rrr(pkg, oracle):
    rng = np.random.default_rng(91)
    for name, t in texts.text_catalogue(large=True):
        n = len(t)
        chk = oracle.wt_huff(t)
        with pkg.WtHuff(t, flags=pkg.F_RRR_BV) as wt:
            assert wt.size == n and wt.sigma == len(set(t)), name
            i, c = texts.wt_queries(t, rng, min(40000, 20 * n + 16))
            assert (wt.rank(i, c) == chk.rank(i, c)).all(), (name, "rank")
            j = rng.integers(0, n, len(i), dtype=np.uint64)
            r, s = wt.inverse_select(j)
            rr, ss = chk.inverse_select(j)
            assert (r == rr).all() and (s == ss).all(), (name, "inverse_select")
            assert (wt.select(r + np.uint64(1), s.astype(np.uint8)) == j).all(), (name, "select round trip")
            tot = wt.rank(np.full(256, n, dtype=np.uint64), np.arange(256, dtype=np.uint8))
            assert (tot == np.bincount(np.frombuffer(t, np.uint8), minlength=256).astype(np.uint64)).all(), name


def test_csa_over_rrr(pkg, oracle):
    rng = np.random.default_rng(92)
    for name, t in texts.text_catalogue(zero_free=True, large=True):
        if len(t) > 1_500_000:
            continue
        chk = oracle.csa(t)
        pats = [t[s : s + int(rng.integers(1, 16))] for s in rng.integers(0, max(1, len(t) - 16), 500)] + [b"", b"\x01\x02zz", t[:3]]
        flat, off = pkg.csr_patterns(pats)
        with pkg.CsaWt(t, flags=pkg.F_RRR_BV) as csa, pkg.CsaWt(t) as plain:
            cnt = csa.count(flat, off)
            assert (cnt == chk.count(flat, off)).all(), (name, "count")
            keep = [p for p, c in zip(pats, cnt) if c <= 3000]
            kflat, koff = pkg.csr_patterns(keep)
            a, b = csa.locate(kflat, koff), chk.locate(kflat, koff)
            assert (a[0] == b[0]).all() and (a[1] == b[1]).all(), (name, "locate")
            k = rng.integers(0, len(t) + 1, 2000, dtype=np.uint64)
            assert (csa.sa(k) == chk.sa(k)).all(), (name, "SA access")
            bq = rng.integers(0, len(t) + 1, 500, dtype=np.uint64)
            eq = np.minimum(bq + rng.integers(0, 50, 500, dtype=np.uint64), np.uint64(len(t)))
            assert (csa.extract(bq, eq)[1] == plain.extract(bq, eq)[1]).all(), (name, "extract")
            assert csa.device_bytes <= plain.device_bytes or len(t) < 100000, (name, "compressed image is not larger")


def test_load_reference_blobs_over_rrr(pkg, orc):
    """ingest of wt_huff<rrr_vector<63>> / csa_wt<wt_huff<rrr_vector<63>>> as serialised by the reference"""
    if not orc.ref_available() or not hasattr(orc.Ref().L, "ref_wt_huff_rrr_create"):
        pytest.skip("reference library without the rrr-backed types")
    ref = orc.Ref()
    rng = np.random.default_rng(93)
    for name, t in texts.text_catalogue(large=False):
        blob, ref_rank = ref.wt_huff_rrr_blob(t)
        with pkg.load_sdsl(blob, pkg.KIND_WT_HUFF, flags=pkg.F_RRR_BV) as wt:
            i, c = texts.wt_queries(t, rng, 5000)
            assert wt.size == len(t) and (wt.rank(i, c) == ref_rank(i, c)).all(), (name, "wt rank")
            j = rng.integers(0, len(t), 3000, dtype=np.uint64)
            r, s = wt.inverse_select(j)
            assert (s == np.frombuffer(t, np.uint8)[j.astype(np.int64)]).all(), (name, "access")
            assert (wt.select(r + np.uint64(1), s.astype(np.uint8)) == j).all(), (name, "select")
    for name, t in texts.text_catalogue(zero_free=True, large=False):
        blob, ref_count = ref.csa_rrr_blob(t)
        pats = [t[s : s + int(rng.integers(1, 12))] for s in rng.integers(0, max(1, len(t) - 12), 300)] + [b"", b"\x01\x02zz"]
        flat, off = pkg.csr_patterns(pats)
        with pkg.load_sdsl(blob, pkg.KIND_CSA_WT, flags=pkg.F_RRR_BV) as csa:
            assert csa.size == len(t) + 1 and (csa.count(flat, off) == ref_count(flat, off)).all(), (name, "count")
            bq = rng.integers(0, len(t), 200, dtype=np.uint64)
            eq = np.minimum(bq + rng.integers(0, 30, 200, dtype=np.uint64), np.uint64(len(t) - 1))
            o, out = csa.extract(bq, eq)
            for k in range(0, 200, 9):
                assert out[int(o[k]) : int(o[k + 1])].tobytes() == t[int(bq[k]) : int(eq[k]) + 1], (name, "extract")
