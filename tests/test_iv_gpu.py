"""The int_vector<w> wire format of the C ABI (sdslgpu_rank_iv / sdslgpu_select_iv): packing helpers on the CPU against
plain integer arithmetic, and on the GPU the packed calls against the u64 calls and the oracle (host and device arrays,
several widths, a batch that spans two staging chunks)."""
import numpy as np
import pytest

import cases


def test_iv_pack_helpers_cpu(pkg):
    rng = np.random.default_rng(1)
    for w in (1, 5, 20, 34, 40, 63, 64):
        for n in (0, 1, 63, 64, 65, 1000):
            v = rng.integers(0, 2**64, n, dtype=np.uint64)
            m = (1 << w) - 1
            pk = pkg.iv_pack(v, w)
            assert len(pk) == pkg.iv_words(n, w)
            big = 0
            for k, x in enumerate(v):
                big |= (int(x) & m) << (k * w)  # int_vector<w>: field k at bit k*w, LSB first (bits.hpp:737-790)
            for j in range(len(pk)):
                assert int(pk[j]) == (big >> (64 * j)) & (2**64 - 1), (w, n, j)
            assert (pkg.iv_unpack(pk, w, n) == (v & np.uint64(m))).all()


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["bv", "rrr", "sd"])
def test_iv_calls_match_u64_calls(pkg, oracle, kind):
    import torch

    nbits = 2_000_003
    w = cases.bernoulli_words(nbits, 0.3, 5)
    ob = oracle.bv(w, nbits)
    cls = {"bv": pkg.BitVector, "rrr": pkg.RrrVector, "sd": pkg.SdVector}[kind]
    rng = np.random.default_rng(3)
    with cls(w, nbits) as v:
        n = 100_001
        idx = rng.integers(0, nbits + 1, n, dtype=np.uint64)
        for b in (1, 0):
            m = v.arg_count(b)
            sel = rng.integers(1, m + 1, n, dtype=np.uint64)
            want_r, want_s = ob.rank(idx, b), ob.select(sel, b)
            for win, wout in ((21, 21), (34, 64), (64, 22), (40, 40)):
                mo = np.uint64((1 << wout) - 1) if wout < 64 else np.uint64(2**64 - 1)
                pi, ps = pkg.iv_pack(idx, win), pkg.iv_pack(sel, win)
                r = pkg.iv_unpack(v.rank_iv(pi, win, n, b, wout), wout, n)
                assert (r == (want_r & mo)).all(), (kind, b, win, wout, "rank host")
                s_ = pkg.iv_unpack(v.select_iv(ps, win, n, b, wout), wout, n)
                assert (s_ == (want_s & mo)).all(), (kind, b, win, wout, "select host")
                d = torch.from_numpy(pi.view(np.int64)).cuda()
                got = v.rank_iv(d, win, n, b, wout)
                torch.cuda.synchronize()
                assert (pkg.iv_unpack(got.cpu().numpy().view(np.uint64), wout, n) == (want_r & mo)).all(), (kind, b, win, wout, "rank device")
        # out of domain -> all-ones field
        bad = np.array([nbits + 1, 0, nbits], dtype=np.uint64)
        r = pkg.iv_unpack(v.rank_iv(pkg.iv_pack(bad, 30), 30, 3, 1, 30), 30, 3)
        assert r[0] == (1 << 30) - 1 and r[1] == 0 and r[2] == ob.rank(np.array([nbits], np.uint64), 1)[0]


@pytest.mark.gpu
def test_iv_batch_spanning_chunks(pkg):
    """2^23 + 1000 queries = two staging chunks; every chunk starts on a word boundary for any width"""
    nbits = 50_000_017
    w = cases.random_words(nbits, 9)
    rng = np.random.default_rng(4)
    n = (1 << 23) + 1000
    idx = rng.integers(0, nbits + 1, n, dtype=np.uint64)
    with pkg.BitVector(w, nbits) as v:
        want = v.rank(idx, 1)
        for width in (26, 27):
            got = pkg.iv_unpack(v.rank_iv(pkg.iv_pack(idx, width), width, n, 1, width), width, n)
            assert (got == want).all(), width
        m = v.arg_count(1)
        sel = rng.integers(1, m + 1, n, dtype=np.uint64)
        got = pkg.iv_unpack(v.select_iv(pkg.iv_pack(sel, 26), 26, n, 1, 26), 26, n)
        assert (got == v.select(sel, 1)).all()
