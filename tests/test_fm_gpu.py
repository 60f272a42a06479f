"""GPU parity: csa_wt<wt_huff<>> count / locate / SA access / bwt.rank through the C ABI (SURVEY.md §8 rows
a9, a10) against the oracle, the unmodified reference, and a brute-force scan (count/locate are unpinned by
the reference's own tests)."""
import numpy as np
import pytest

import texts

pytestmark = pytest.mark.gpu


def _patterns(t, rng, k, maxlen=24):
    n = len(t)
    pats = []
    for _ in range(k):
        m = int(rng.integers(1, maxlen + 1))
        if n and rng.random() < 0.7:
            s = int(rng.integers(0, max(1, n - m + 1)))
            pats.append(t[s : s + m])
        else:
            pats.append(rng.integers(1, 256, m, dtype=np.uint8).tobytes())
    return pats + [b"", t, t[:4], t + b"x"]


# the searches run on the one-hot occurrence bitmaps by default and on the wavelet tree alone with F_COMPACT
VARIANTS = [("occ16", 0), ("compact", 8)]


@pytest.mark.parametrize("variant,flags", VARIANTS)
def test_fm_catalogue(pkg, oracle, orc, variant, flags):
    rng = np.random.default_rng(21)
    assert pkg.F_COMPACT == 8
    for name, t in texts.text_catalogue(zero_free=True, large=True):
        with pkg.CsaWt(t, flags=flags) as csa:
            assert csa.size == len(t) + 1, name
            pats = _patterns(t, rng, 1500)
            flat, off = pkg.csr_patterns(pats)
            cnt, l = csa.count(flat, off, want_l=True)
            small = [p for p, c in zip(pats, cnt) if c <= 5000]
            sflat, soff = pkg.csr_patterns(small)
            occ_off, occ = csa.locate(sflat, soff)
            i = rng.integers(0, len(t) + 1, 4000, dtype=np.uint64)
            sa = csa.sa(i)
            qi, qc = texts.wt_queries(t + b"\0", rng, 4000)
            br = csa.bwt_rank(qi, qc)
            checkers = [("oracle", oracle.csa(t))]
            if orc.ref_available():
                checkers.append(("reference", orc.Ref().csa(t)))
            for cname, chk in checkers:
                c2, l2 = chk.count(flat, off, want_l=True)
                assert (cnt == c2).all(), (name, cname, "count")
                assert (l[cnt > 0] == l2[cnt > 0]).all(), (name, cname, "interval")
                o2 = chk.locate(sflat, soff)
                assert (occ_off == o2[0]).all() and (occ == o2[1]).all(), (name, cname, "locate order")
                assert (sa == chk.sa(i)).all(), (name, cname, "SA access")
                if cname == "reference":
                    assert (br == chk.bwt_rank(qi, qc)).all(), (name, "bwt.rank")
            # brute force on a few patterns
            for k in rng.integers(0, len(small), 25):
                p = small[int(k)]
                if not p or len(t) > 200000:
                    continue
                want, s = [], t.find(p)
                while s >= 0:
                    want.append(s)
                    s = t.find(p, s + 1)
                assert sorted(occ[int(occ_off[k]) : int(occ_off[k + 1])].tolist()) == want, (name, p)


def test_fm_rejects_zero_byte(pkg):
    with pytest.raises(pkg.SdslGpuError):
        pkg.CsaWt(b"abc\0def")


def test_fm_device_buffers(pkg, oracle):
    import torch

    rng = np.random.default_rng(2)
    t = rng.integers(1, 256, 2_000_000, dtype=np.uint8).tobytes()
    pats = [t[s : s + 20] for s in rng.integers(0, len(t) - 20, 50000)]
    flat, off = pkg.csr_patterns(pats)
    with pkg.CsaWt(t) as csa:
        d_flat, d_off = torch.from_numpy(flat).cuda(), torch.from_numpy(off.view(np.int64)).cuda()
        cnt = csa.count(d_flat, d_off)
        torch.cuda.synchronize()
        host = csa.count(flat, off)
        assert (cnt.cpu().numpy().view(np.uint64) == host).all() and (host >= 1).all()
        occ_off, occ = csa.locate(d_flat, d_off)
        torch.cuda.synchronize()
        occ_off, occ = occ_off.cpu().numpy().view(np.uint64), occ.cpu().numpy().view(np.uint64)
        # every reported position really is an occurrence; patterns sampled at s are found at s
        arr = np.frombuffer(t, dtype=np.uint8)
        for k in rng.integers(0, len(pats), 300):
            for p in occ[int(occ_off[k]) : int(occ_off[k + 1])]:
                assert t[int(p) : int(p) + 20] == pats[int(k)]
        chk = oracle.csa(t)
        assert (host == chk.count(flat, off)).all()


def test_gpu_suffix_array_matches_host_sais(pkg, monkeypatch):
    """construction: the device prefix-doubling suffix sorter and the host SA-IS builder give the same index
    (same counts, same locate order, same SA values) on random, repetitive and natural-language-like texts"""
    rng = np.random.default_rng(31)
    for name, t in texts.text_catalogue(zero_free=True, large=True):
        pats = _patterns(t, rng, 300)
        flat, off = pkg.csr_patterns(pats)
        i = rng.integers(0, len(t) + 1, 3000, dtype=np.uint64)
        monkeypatch.setenv("SDSLGPU_HOST_SA", "1")
        with pkg.CsaWt(t) as a:
            want = (a.count(flat, off), a.sa(i))
        monkeypatch.delenv("SDSLGPU_HOST_SA")
        with pkg.CsaWt(t) as b:
            assert (b.count(flat, off) == want[0]).all() and (b.sa(i) == want[1]).all(), name


@pytest.mark.parametrize("variant,flags", VARIANTS)
def test_extract(pkg, oracle, orc, variant, flags):
    """sdsl::extract (suffix_array_algorithm.hpp:590-610) == the text itself, == the oracle / reference"""
    rng = np.random.default_rng(41)
    for name, t in texts.text_catalogue(zero_free=True, large=(variant == "occ16")):
        n = len(t) + 1
        full = t + b"\0"
        b = rng.integers(0, n, 3000, dtype=np.uint64)
        e = np.minimum(b + rng.integers(0, 100, 3000, dtype=np.uint64), np.uint64(n - 1))
        b[0], e[0] = 0, min(n - 1, 5000)
        with pkg.CsaWt(t, flags=flags) as csa:
            off, out = csa.extract(b, e)
            assert off[-1] == int((e - b + 1).sum())
            for k in range(0, len(b), 7):
                assert out[int(off[k]) : int(off[k + 1])].tobytes() == full[int(b[k]) : int(e[k]) + 1], (name, k)
            if len(t) <= 200000:
                o_off, o_out = oracle.csa(t).extract(b, e)
                assert (o_off == off).all() and (o_out == out).all(), (name, "oracle")
                if orc.ref_available():
                    rc = orc.Ref().csa(t)
                    for k in range(0, 40):
                        assert rc.extract(int(b[k]), int(e[k])) == out[int(off[k]) : int(off[k + 1])].tobytes(), (name, "reference")
            # an index ingested from the reference's serialised bytes extracts the same text
            if len(t) <= 200000 and orc.ref_available():
                with pkg.load_sdsl(orc.Ref().csa(t).serialize(), pkg.KIND_CSA_WT, flags=flags) as loaded:
                    l_off, l_out = loaded.extract(b, e)
                    assert (l_out == out).all(), (name, "loaded blob")


@pytest.mark.parametrize("sa_dens,isa_dens", [(1, 1), (3, 5), (8, 16), (64, 128)])
def test_sampling_densities_do_not_change_results(pkg, oracle, sa_dens, isa_dens):
    """t_dens / t_inv_dens (csa_wt.hpp:50-51) trade memory for LF steps; count / locate / csa[i] / extract stay
    those of the default-density oracle"""
    rng = np.random.default_rng(43)
    for name, t in texts.text_catalogue(zero_free=True):
        n = len(t) + 1
        pats = _patterns(t, rng, 200)
        flat, off = pkg.csr_patterns(pats)
        i = rng.integers(0, n, 2000, dtype=np.uint64)
        b = rng.integers(0, n, 500, dtype=np.uint64)
        e = np.minimum(b + rng.integers(0, 60, 500, dtype=np.uint64), np.uint64(n - 1))
        chk = oracle.csa(t)
        with pkg.CsaWt(t, sa_dens=sa_dens, isa_dens=isa_dens, flags=pkg.F_COMPACT if sa_dens == 3 else 0) as csa:
            assert (csa.count(flat, off) == chk.count(flat, off)).all(), name
            a, w = csa.locate(flat, off), chk.locate(flat, off)
            assert (a[0] == w[0]).all() and (a[1] == w[1]).all(), name
            assert (csa.sa(i) == chk.sa(i)).all(), name
            got, want = csa.extract(b, e), chk.extract(b, e)
            assert (got[0] == want[0]).all() and (got[1] == want[1]).all(), name


@pytest.mark.parametrize("distinct", [1, 15, 16, 17, 32, 33, 255])
def test_occ16_alphabet_edges(pkg, oracle, distinct):
    """the occurrence bitmaps switch from one level to two at sigma = 17 (text symbols + the sentinel): both sides of
    the switch, a level-0 bitmap with a single symbol, and the full byte alphabet"""
    rng = np.random.default_rng(100 + distinct)
    t = (rng.integers(0, distinct, 40000, dtype=np.uint8) + 1).astype(np.uint8).tobytes()
    n = len(t) + 1
    pats = _patterns(t, rng, 400, maxlen=6)
    flat, off = pkg.csr_patterns(pats)
    i = rng.integers(0, n, 3000, dtype=np.uint64)
    b = rng.integers(0, n, 300, dtype=np.uint64)
    e = np.minimum(b + rng.integers(0, 50, 300, dtype=np.uint64), np.uint64(n - 1))
    chk = oracle.csa(t)
    want = (chk.count(flat, off, want_l=True), chk.sa(i), chk.extract(b, e))
    for flags in (0, pkg.F_COMPACT):
        with pkg.CsaWt(t, flags=flags) as csa:
            cnt, l = csa.count(flat, off, want_l=True)
            assert (cnt == want[0][0]).all() and (l[cnt > 0] == want[0][1][cnt > 0]).all(), (distinct, flags)
            assert (csa.sa(i) == want[1]).all(), (distinct, flags)
            got = csa.extract(b, e)
            assert (got[0] == want[2][0]).all() and (got[1] == want[2][1]).all(), (distinct, flags)
    # the default index pays for its speed in device memory; the compact one is the wavelet tree alone
    with pkg.CsaWt(t) as a, pkg.CsaWt(t, flags=pkg.F_COMPACT) as c:
        assert a.device_bytes > c.device_bytes
