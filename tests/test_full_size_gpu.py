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
