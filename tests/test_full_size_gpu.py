"""GPU, at BASELINE.json's FULL sizes, through size-independent properties (the oracle cannot run these sizes in
seconds): config 2 (2^33-bit vector, rank + select), config 3 (rrr / sd on 2^33 bits), config 4 (wt_huff on 2^28
bytes), config 5 (csa_wt on a 2^30-byte text: count / locate / extract checked against the text itself)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rand_words(nw, seed):
    import torch

    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randint(-(2**63), 2**63 - 1, (nw,), dtype=torch.int64, device="cuda", generator=g)


def test_config2_and_3_full_size(pkg):
    import torch

    nbits = 1 << 33
    words = _rand_words(nbits // 64, 42)
    g = torch.Generator(device="cuda").manual_seed(7)
    idx = torch.randint(0, nbits + 1, (20_000_000,), dtype=torch.int64, device="cuda", generator=g)
    with pkg.BitVector(words, nbits) as bv:
        m1 = bv.arg_count(1)
        assert abs(m1 - nbits // 2) < 1 << 20  # ~2^32 ones
        r1, r0 = bv.rank(idx, 1), bv.rank(idx, 0)
        assert bool((r1 + r0 == idx).all())
        # rank against a direct popcount of the words below idx for a subsample (bit-exact, no oracle needed)
        sub = idx[:2000].cpu().numpy().astype(np.uint64)
        wh = words[: int(sub.max() // 64) + 2].cpu().numpy().view(np.uint64)
        for k in range(0, 2000, 97):
            i = int(sub[k])
            full = int(np.unpackbits(wh[: i // 64].view(np.uint8)).sum()) if i >= 64 else 0
            part = bin(int(wh[i // 64]) & ((1 << (i % 64)) - 1)).count("1")
            assert int(r1[k]) == full + part
        for b, m in ((1, m1), (0, nbits - m1)):
            k = torch.randint(1, m + 1, (20_000_000,), dtype=torch.int64, device="cuda", generator=g)
            p = bv.select(k, b)
            assert bool((bv.rank(p, b) == k - 1).all()) and bool((bv.access(p) == b).all())
        with pkg.RrrVector(words, nbits) as rrr, pkg.SdVector(words, nbits) as sd:
            q = idx[:5_000_000]
            assert bool((rrr.rank(q, 1) == r1[:5_000_000]).all()) and bool((sd.rank(q, 1) == r1[:5_000_000]).all())
            k = torch.randint(1, m1 + 1, (5_000_000,), dtype=torch.int64, device="cuda", generator=g)
            want = bv.select(k, 1)
            assert bool((rrr.select(k, 1) == want).all()) and bool((sd.select(k, 1) == want).all())


def _bernoulli_words_gpu(nbits, density, seed):
    import torch

    g = torch.Generator(device="cuda").manual_seed(seed)
    nw = nbits // 64
    out = torch.empty(nw, dtype=torch.int64, device="cuda")
    w = torch.ones(64, dtype=torch.int64, device="cuda") << torch.arange(64, device="cuda", dtype=torch.int64)
    chunk = 1 << 21
    for lo in range(0, nw, chunk):
        hi = min(nw, lo + chunk)
        out[lo:hi] = ((torch.rand((hi - lo, 64), device="cuda", generator=g) < density).to(torch.int64) * w).sum(1)
    return out


def test_config2_and_3_full_size_against_the_reference(pkg, orc):
    """BASELINE configs 2 / 3 at their full size (2^33 bits), every density of the sweep (1, 5, 10, 25, 50 %): a sample
    of 2e5 rank, select_1 and select_0 answers of the plain vector, rrr_vector<63> and sd_vector<> compared with the
    UNMODIFIED reference built from the same bits on the host (the ten reference constructions run in parallel host
    threads while the GPU structures are built and queried)."""
    import concurrent.futures as cf

    import torch

    if not orc.ref_available():
        pytest.skip("needs oracle/_ref/libsdslref.so (the unmodified reference)")
    R = orc.Ref()
    nbits, ns = 1 << 33, 200_000
    densities = (0.01, 0.05, 0.10, 0.25, 0.50)
    rng = np.random.default_rng(11)
    idx = rng.integers(0, nbits + 1, ns, dtype=np.uint64)
    idx[:2] = [0, nbits]
    with cf.ThreadPoolExecutor(max_workers=10) as pool:
        jobs = {}
        host_words = {}
        for d in densities:
            wd = _bernoulli_words_gpu(nbits, d, 1000 + int(d * 100))
            wh = wd.cpu().numpy().view(np.uint64)
            del wd
            host_words[d] = wh
            for kind in ("bv", "rrr", "sd"):  # ctypes releases the GIL: the reference builds overlap each other and the GPU work
                jobs[(d, kind)] = pool.submit((lambda k, w: getattr(R, k)(w, nbits)), kind, wh)
        for d in densities:
            wh = host_words[d]
            wt_ = torch.from_numpy(wh.view(np.int64)).cuda()
            with pkg.BitVector(wt_, nbits) as bv, pkg.RrrVector(wt_, nbits) as rrr, pkg.SdVector(wt_, nbits) as sd:
                del wt_
                sel = {b: rng.integers(1, bv.arg_count(b) + 1, ns, dtype=np.uint64) for b in (0, 1)}
                got = {}
                for name, v in (("bv", bv), ("rrr", rrr), ("sd", sd)):
                    got[name] = {"rank1": v.rank(idx, 1), "rank0": v.rank(idx, 0), "select1": v.select(sel[1], 1), "select0": v.select(sel[0], 0)}
                for name in ("bv", "rrr", "sd"):
                    ref = jobs[(d, name)].result()
                    assert (got[name]["rank1"] == ref.rank(idx, 1, threads=8)).all(), (d, name, "rank_1")
                    assert (got[name]["rank0"] == ref.rank(idx, 0, threads=8)).all(), (d, name, "rank_0")
                    assert (got[name]["select1"] == ref.select(sel[1], 1, threads=8)).all(), (d, name, "select_1")
                    assert (got[name]["select0"] == ref.select(sel[0], 0, threads=8)).all(), (d, name, "select_0")
                    del ref
                    jobs[(d, name)] = None
            host_words[d] = None


def test_config4_full_size(pkg):
    rng = np.random.default_rng(42)
    n = 1 << 28
    text = rng.integers(0, 256, n, dtype=np.uint8)
    with pkg.WtHuff(text) as wt:
        tot = wt.rank(np.full(256, n, dtype=np.uint64), np.arange(256, dtype=np.uint8))
        assert (tot == np.bincount(text, minlength=256).astype(np.uint64)).all()
        j = rng.integers(0, n, 10_000_000, dtype=np.uint64)
        rnk, sym = wt.inverse_select(j)
        assert (sym == text[j.astype(np.int64)]).all()
        c = sym.astype(np.uint8)
        assert (wt.rank(j, c) == rnk).all() and (wt.rank(j + np.uint64(1), c) == rnk + np.uint64(1)).all()
        assert (wt.select(rnk + np.uint64(1), c) == j).all()


def test_config4_full_size_against_the_reference(pkg, orc):
    """C4 (wt_huff<> on 2^28 bytes): 2e5 rank(i, c), select(i, c), inverse_select(i) answers against the reference's own
    wt_huff built from the same text on the host"""
    if not orc.ref_available():
        pytest.skip("needs oracle/_ref/libsdslref.so (the unmodified reference)")
    rng = np.random.default_rng(42)
    n, ns = 1 << 28, 200_000
    text = rng.integers(0, 256, n, dtype=np.uint8)
    ref = orc.Ref().wt_huff(text)
    with pkg.WtHuff(text) as wt:
        i = rng.integers(0, n + 1, ns, dtype=np.uint64)
        c = rng.integers(0, 256, ns, dtype=np.uint8)
        assert (wt.rank(i, c) == ref.rank(i, c, threads=8)).all()
        occ = wt.rank(np.full(256, n, dtype=np.uint64), np.arange(256, dtype=np.uint8))
        k = (rng.integers(0, 2**62, ns, dtype=np.uint64) % occ[c.astype(np.int64)]) + np.uint64(1)
        assert (wt.select(k, c) == ref.select(k, c, threads=8)).all()
        j = rng.integers(0, n, ns, dtype=np.uint64)
        rnk, sym = wt.inverse_select(j)
        rr, rs = ref.inverse_select(j, threads=8)
        assert (rnk == rr).all() and (sym == rs).all()


def test_config5_full_size(pkg):
    rng = np.random.default_rng(42)
    n = 1 << 30
    text = rng.integers(1, 256, n, dtype=np.uint8)
    npat, plen = 1_000_000, 20
    starts = rng.integers(0, n - plen, npat)
    flat = text[(starts[:, None] + np.arange(plen)[None, :])].reshape(-1).copy()
    off = np.arange(npat + 1, dtype=np.uint64) * np.uint64(plen)
    with pkg.CsaWt(text) as csa:
        assert csa.size == n + 1
        cnt = csa.count(flat, off)
        assert (cnt >= 1).all() and cnt.sum() < npat + 100  # 20-mers of a uniform text are unique
        occ_off, occ = csa.locate(flat, off)
        assert occ_off[-1] == cnt.sum()
        # every reported position spells the pattern; the sampled start itself is among them
        first = occ[occ_off[:-1].astype(np.int64)].astype(np.int64)
        got = text[(first[:200000, None] + np.arange(plen)[None, :])]
        assert (got.reshape(-1) == flat[: 200000 * plen]).all()
        single = cnt == 1
        assert (occ[occ_off[:-1][single].astype(np.int64)] == starts[single].astype(np.uint64)).all()
        # absent patterns: a byte that is not in the text, and random 20-mers
        rflat = rng.integers(1, 256, 1000 * plen, dtype=np.uint8)
        assert csa.count(rflat, off[:1001]).sum() <= 2
        b = rng.integers(0, n - 64, 20000).astype(np.uint64)
        o, out = csa.extract(b, b + np.uint64(63))
        assert (out.reshape(-1, 64) == text[(b.astype(np.int64)[:, None] + np.arange(64)[None, :])]).all()
        # ... and against the reference: it loads the index this engine serialises (csa_wt::serialize bytes, 1.9 GB)
        # and answers count / locate for 1e4 of the patterns (plus absent ones) itself
        orc = __import__("__graft_entry__").load_oracle()
        if orc.ref_available():
            ref = orc.Ref().csa(blob=csa.serialize(0))
            assert ref.size == n + 1
            k = 10_000
            mixed = np.concatenate([flat[: k * plen], rflat[: 100 * plen]])
            moff = np.arange(k + 100 + 1, dtype=np.uint64) * np.uint64(plen)
            assert (csa.count(mixed, moff) == ref.count(mixed, moff, threads=8)).all()
            go, gc = csa.locate(mixed, moff)
            ro, rc = ref.locate(mixed, moff, threads=8)
            assert (go == ro).all() and (gc == rc).all()
