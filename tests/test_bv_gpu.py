"""GPU parity: batched rank / select / access on plain bit vectors through the C ABI, against the oracle
(oracle/oracle.c) and, where built, the unmodified reference (oracle/_ref) — SURVEY.md §8 rows a2-a4."""
import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu


def _checkers(oracle, orc, w, nbits):
    out = [("oracle", oracle.bv(w, nbits))]
    if orc.ref_available():
        out.append(("reference", orc.Ref().bv(w, nbits)))
    return out


@pytest.mark.parametrize("flags", [0, 1], ids=["b200_layout", "sdsl_layout"])
def test_rank_select_catalogue(pkg, oracle, orc, flags):
    """edge-case catalogue of test/rank_support_test.config / select_support_test.config: every position
    of the small vectors, 60k random positions of the 1e6-bit ones, both patterns"""
    for cid, w, nbits in cases.bitvector_catalogue(large=True):
        with pkg.BitVector(w, nbits, flags=flags) as bv:
            assert bv.size == nbits
            idx = cases.rank_queries(nbits, 11, 60000)
            for name, chk in _checkers(oracle, orc, w, nbits):
                for b in (1, 0):
                    assert (bv.rank(idx, b) == chk.rank(idx, b)).all(), (cid, name, "rank", b)
                    m = bv.arg_count(b)
                    assert m == int(chk.rank([nbits], b)[0]), (cid, name, "arg_count", b)
                    q = cases.select_queries(m, 12, 60000)
                    if len(q):
                        assert (bv.select(q, b) == chk.select(q, b)).all(), (cid, name, "select", b)
            if nbits:
                bits = cases.unpack_bits(w, nbits)
                pos = idx[idx < nbits]
                assert (bv.access(pos) == bits[pos.astype(np.int64)]).all(), (cid, "access")


def test_sdsl_tables_built_on_device_are_byte_identical(pkg, oracle, orc):
    """construction parity: m_basic_block built by sdsl_table_*_kernel == rank_support_v::serialize bytes"""
    for cid, w, nbits in cases.bitvector_catalogue(large=True):
        with pkg.BitVector(w, nbits, flags=pkg.F_SDSL_LAYOUT) as bv:
            ob = oracle.bv(w, nbits)
            rb = orc.Ref().bv(w, nbits, with_select=False) if orc.ref_available() else None
            for what in (0, 1, 2):
                got = bv.serialize(what)
                assert got == ob.serialize(what), (cid, what, "oracle")
                if rb is not None:
                    assert got == rb.serialize(what), (cid, what, "reference")


def test_out_of_domain_is_defined(pkg):
    w = cases.random_words(1000, 5)
    with pkg.BitVector(w, 1000) as bv:
        r = bv.rank(np.array([1000, 1001, 2**63], dtype=np.uint64))
        assert r[0] == bv.arg_count(1) and r[1] == pkg.NPOS and r[2] == pkg.NPOS
        m = bv.arg_count(1)
        s = bv.select(np.array([0, m, m + 1], dtype=np.uint64))
        assert s[0] == pkg.NPOS and s[2] == pkg.NPOS and s[1] < 1000
        assert bv.access(np.array([1000], dtype=np.uint64))[0] == pkg.NPOS
        assert len(bv.rank(np.zeros(0, np.uint64))) == 0


def test_device_buffers_and_streams(pkg, oracle):
    """device-resident queries (torch tensors) on a non-default stream give the same answers as host buffers"""
    import torch

    nbits = 3_000_001
    w = cases.random_words(nbits, 21, dirty_tail=True)
    idx = cases.rank_queries(nbits, 4, 200000)
    want = oracle.bv(w, nbits).rank(idx, 1)
    with pkg.BitVector(torch.from_numpy(w.view(np.int64)).cuda(), nbits) as bv:
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            d_idx = torch.from_numpy(idx.view(np.int64)).cuda()
            got = bv.rank(d_idx, 1)
        s.synchronize()
        assert (got.cpu().numpy().view(np.uint64) == want).all()
        # host path with more than one staging chunk
        big = np.tile(idx, 25)  # 5e6 queries > 4 Mi chunk
        assert (bv.rank(big, 1) == np.tile(want, 25)).all()


def test_density_sweep_properties(pkg, oracle):
    """size-independent properties on 2^27-bit vectors over the density sweep of BASELINE config 3:
    select(rank(i)+1) >= i, rank(select(k)) == k-1, rank0 + rank1 == idx; spot-checked against the oracle"""
    nbits = (1 << 27) + 12345
    for d in (0.01, 0.1, 0.5, 0.9):
        w = cases.bernoulli_words(nbits, d, int(d * 1000))
        with pkg.BitVector(w, nbits) as bv:
            idx = cases.rank_queries(nbits, 8, 500000)
            r1, r0 = bv.rank(idx, 1), bv.rank(idx, 0)
            assert (r1 + r0 == idx).all()
            for b, r in ((1, r1), (0, r0)):
                m = bv.arg_count(b)
                k = cases.select_queries(m, 9, 500000)
                p = bv.select(k, b)
                assert (bv.rank(p, b) == k - 1).all() and (bv.access(p) == b).all()
                ok = r < m
                nxt = bv.select(r[ok] + 1, b)
                assert (nxt >= idx[ok]).all()
            o = oracle.bv(w, nbits)
            sub = idx[:20000]
            assert (bv.rank(sub, 1) == o.rank(sub, 1)).all()
            ks = cases.select_queries(bv.arg_count(1), 10, 20000)
            assert (bv.select(ks, 1) == o.select(ks, 1)).all()


@pytest.mark.parametrize("log_s", ["0", "3", "8", "10", "12"])
def test_select_sparse_samples_and_interpolation(pkg, oracle, monkeypatch, log_s):
    """the select paths large vectors use (sample stride > 64, interpolated first probe, bisect / walk-left /
    walk-right repairs) forced onto the small catalogue — clustered and skewed vectors included — and compared
    exhaustively with the oracle"""
    monkeypatch.setenv("SDSLGPU_SELECT_LOG_S", log_s)
    monkeypatch.setenv("SDSLGPU_SELECT_INTERP", "1")
    for cid, w, nbits in cases.bitvector_catalogue(large=True):
        o = oracle.bv(w, nbits)
        with pkg.BitVector(w, nbits) as bv:
            for b in (1, 0):
                q = cases.select_queries(bv.arg_count(b), 13, 60000)
                if len(q):
                    assert (bv.select(q, b) == o.select(q, b)).all(), (cid, b, log_s)


def test_select_clustered_large(pkg):
    """2^31 bits, strongly clustered (long empty stretches, dense bursts): automatic stride + interpolation;
    checked through rank(select(k)) == k-1 and bit(select(k)) == b for 2e6 queries per pattern"""
    import torch

    nbits = 1 << 31
    nw = nbits // 64
    g = torch.Generator(device="cuda").manual_seed(3)
    words = torch.zeros(nw, dtype=torch.int64, device="cuda")
    # bursts: 1/16 of the 64-Kbit regions are 90 % dense, the rest nearly empty
    region = torch.rand(nw // 1024, device="cuda", generator=g) < (1 / 16)
    dense = region.repeat_interleave(1024)
    r = torch.randint(-(2**63), 2**63 - 1, (nw,), dtype=torch.int64, device="cuda", generator=g)
    r2 = torch.randint(-(2**63), 2**63 - 1, (nw,), dtype=torch.int64, device="cuda", generator=g)
    r3 = torch.randint(-(2**63), 2**63 - 1, (nw,), dtype=torch.int64, device="cuda", generator=g)
    words = torch.where(dense, r | r2 | r3, r & r2 & r3 & torch.roll(r, 1) & torch.roll(r2, 1) & torch.roll(r3, 2))
    with pkg.BitVector(words, nbits) as bv:
        for b in (1, 0):
            m = bv.arg_count(b)
            k = torch.from_numpy(cases.select_queries(m, 17, 2_000_000).view(np.int64)).cuda()
            p = bv.select(k, b)
            assert bool((bv.rank(p, b) == k - 1).all()) and bool((bv.access(p) == b).all()), b


@pytest.mark.parametrize("flags", [0, 1], ids=["b200_layout", "sdsl_layout"])
def test_two_bit_patterns(pkg, oracle, orc, flags):
    """rank_support_v<10|01|00|11, 2> / select_support_mcl<..., 2> (rank_support_test.cpp:44-47,
    select_support_test.cpp:39-42) over the catalogue: oracle, reference, and dirty tail bits past size()"""
    assert (pkg.PAT_10, pkg.PAT_01, pkg.PAT_00, pkg.PAT_11) == (2, 3, 4, 5)
    for cid, w, nbits in cases.bitvector_catalogue(large=True):
        with pkg.BitVector(w, nbits, flags=flags) as bv:
            idx = cases.rank_queries(nbits, 13, 40000)
            for name, chk in _checkers(oracle, orc, w, nbits):
                for b in (2, 3, 4, 5):
                    assert (bv.rank(idx, b) == chk.rank(idx, b)).all(), (cid, name, "rank", b)
                    m = bv.arg_count(b)
                    assert m == int(chk.rank([nbits], b)[0]), (cid, name, "arg_count", b)
                    q = cases.select_queries(m, 14, 40000)
                    if len(q):
                        assert (bv.select(q, b) == chk.select(q, b)).all(), (cid, name, "select", b)
                    # out of domain: defined here, UB in the reference
                    assert bv.rank(np.array([nbits + 1], np.uint64), b)[0] == pkg.NPOS
                    assert (bv.select(np.array([0, m + 1], np.uint64), b) == pkg.NPOS).all()
    # a compressed vector has no two-bit supports in the reference either
    w, nbits = cases.random_words(5000, 3), 5000
    with pkg.RrrVector(w, nbits) as r:
        with pytest.raises(pkg.SdslGpuError):
            r.rank(np.array([1], np.uint64), 2)


def test_two_bit_patterns_large_properties(pkg):
    """2^30-bit vector: select(rank(i)+1) >= i, rank(select(k)+1) == k, and the four pattern counts add up to n-1"""
    import torch

    nbits = 1 << 30
    g = torch.Generator(device="cuda").manual_seed(5)
    words = torch.randint(-(2**63), 2**63 - 1, (nbits // 64,), dtype=torch.int64, device="cuda", generator=g)
    with pkg.BitVector(words, nbits) as bv:
        total = 0
        k = torch.randint(1, 1 << 27, (2_000_000,), dtype=torch.int64, device="cuda", generator=g)
        for b in (2, 3, 4, 5):
            m = bv.arg_count(b)
            total += m
            assert abs(m - nbits / 4) < 1e6
            pos = bv.select(k, b)
            assert bool((bv.rank(pos + 1, b) == k).all()) and bool((bv.rank(pos, b) == k - 1).all())
        assert total == nbits - 1


@pytest.mark.parametrize("chunk_bytes,sectors", [("64", "1"), ("4096", "1"), ("1048576", "1"), ("4096", "0")])
def test_binned_order_matches_direct_and_oracle(pkg, oracle, monkeypatch, chunk_bytes, sectors):
    """ORDER_BINNED (binned.cu: tile counting sort -> bin-major gathers -> un-sort) forced onto the catalogue with
    tiny bins (many bins, empty bins, bins of one block): same answers as the oracle for both patterns, for batches
    shorter than a tile, of exactly one tile, ragged, and with out-of-domain queries mixed in.  sectors = 1: select runs
    through the select sectors wherever the density allows them (bv_device.cuh), 0: through the samples everywhere"""
    monkeypatch.setenv("SDSLGPU_BIN_CHUNK_BYTES", chunk_bytes)
    monkeypatch.setenv("SDSLGPU_SELECT_SECTORS", sectors)
    for cid, w, nbits in cases.bitvector_catalogue(large=True):
        o = oracle.bv(w, nbits)
        with pkg.BitVector(w, nbits) as bv:
            bv.set_batch_order(pkg.ORDER_BINNED)
            for nq in (1, 777, 8192, 8193, 50001):
                idx = cases.rank_queries(nbits, 31 + nq, nq)
                bad = np.arange(len(idx)) % 97 == 5
                idx_bad = idx.copy()
                idx_bad[bad] = np.uint64(nbits + 1) + (idx[bad] << np.uint64(20))
                for b in (1, 0):
                    want = o.rank(idx, b)
                    assert (bv.rank(idx, b) == want).all(), (cid, "rank", b, nq)
                    want_bad = want.copy()
                    want_bad[bad] = pkg.NPOS
                    assert (bv.rank(idx_bad, b) == want_bad).all(), (cid, "rank+ood", b, nq)
                    m = bv.arg_count(b)
                    q = cases.select_queries(m, 32 + nq, nq)
                    if len(q):
                        wsel = o.select(q, b)
                        assert (bv.select(q, b) == wsel).all(), (cid, "select", b, nq)
                        qb = q.copy()
                        badq = np.arange(len(q)) % 89 == 3
                        qb[badq] = np.where(np.arange(len(q))[badq] % 2 == 0, 0, m + 1).astype(np.uint64)
                        wb = wsel.copy()
                        wb[badq] = pkg.NPOS
                        assert (bv.select(qb, b) == wb).all(), (cid, "select+ood", b, nq)
                    else:
                        assert (bv.select(np.array([0, 1, 5], np.uint64), b) == pkg.NPOS).all()
            # two-bit patterns ride on the same launchers
            idx = cases.rank_queries(nbits, 5, 20000)
            assert (bv.rank(idx, pkg.PAT_10) == o.rank(idx, pkg.PAT_10)).all(), (cid, "rank10")


def test_binned_order_large_device_batches(pkg):
    """2^31-bit vector (index > L2), 6e6 device-resident queries: AUTO picks the binned path; BINNED == DIRECT bit
    for bit on uniform and on fully skewed batches (every query in one bin), both patterns, on a side stream"""
    import torch

    nbits = (1 << 31) + 4321
    g = torch.Generator(device="cuda").manual_seed(9)
    words = torch.randint(-(2**63), 2**63 - 1, ((nbits + 63) // 64,), dtype=torch.int64, device="cuda", generator=g)
    nq = 6_000_000 + 17
    with pkg.BitVector(words, nbits) as bv:
        uni = torch.randint(0, nbits + 1, (nq,), dtype=torch.int64, device="cuda", generator=g)
        skew = torch.randint(12345, 12345 + 5000, (nq,), dtype=torch.int64, device="cuda", generator=g)
        mixed = uni.clone()
        mixed[::3] = nbits + 7  # out of domain
        st = torch.cuda.Stream()
        for b in (1, 0):
            m = bv.arg_count(b)
            sel = torch.randint(1, m + 1, (nq,), dtype=torch.int64, device="cuda", generator=g)
            sel_skew = torch.randint(m - 3000, m + 1, (nq,), dtype=torch.int64, device="cuda", generator=g)
            res = {}
            for order in (pkg.ORDER_DIRECT, pkg.ORDER_BINNED, pkg.ORDER_AUTO):
                bv.set_batch_order(order)
                torch.cuda.synchronize()
                with torch.cuda.stream(st):
                    res[order] = [bv.rank(x, b) for x in (uni, skew, mixed)] + [bv.select(x, b) for x in (sel, sel_skew)]
                st.synchronize()
            for k in range(5):
                assert bool((res[pkg.ORDER_DIRECT][k] == res[pkg.ORDER_BINNED][k]).all()), (b, k, "binned")
                assert bool((res[pkg.ORDER_DIRECT][k] == res[pkg.ORDER_AUTO][k]).all()), (b, k, "auto")
            assert bool((res[pkg.ORDER_BINNED][2][::3] == -1).all())
            p = res[pkg.ORDER_BINNED][3]
            bv.set_batch_order(pkg.ORDER_DIRECT)
            assert bool((bv.rank(p, b) == sel - 1).all())


def test_select_sectors_are_built_once_and_only_where_they_apply(pkg, oracle):
    """Select sectors (include/sdslgpu.h, memory note of sdslgpu_select): built by the first select batch that runs
    through the pipeline, per bit value, counted by device_bytes; not for SDSLGPU_F_COMPACT handles, not for sparse
    vectors, not by direct-order batches; answers are the same with and without them, in both batch orders."""
    rng = np.random.default_rng(41)
    nbits = 3_000_017
    dense = cases.pack_bits((rng.random(nbits) < 0.5).astype(np.uint8))
    sparse = cases.pack_bits((rng.random(nbits) < 0.02).astype(np.uint8))
    o = oracle.bv(dense, nbits)
    with pkg.BitVector(dense, nbits) as bv, pkg.BitVector(dense, nbits, flags=pkg.F_COMPACT) as compact, pkg.BitVector(sparse, nbits) as sp:
        b0 = bv.device_bytes
        q1 = cases.select_queries(bv.arg_count(1), 3, 40000)
        q0 = cases.select_queries(bv.arg_count(0), 4, 40000)
        want1, want0 = o.select(q1, 1), o.select(q0, 0)
        bv.set_batch_order(pkg.ORDER_DIRECT)
        assert (bv.select(q1, 1) == want1).all() and bv.device_bytes == b0  # a direct batch builds nothing
        bv.set_batch_order(pkg.ORDER_BINNED)
        assert (bv.select(q1, 1) == want1).all()
        b1 = bv.device_bytes
        assert b1 > b0 + bv.arg_count(1) // 81 * 32 * 9 // 10  # ~32 bytes per 81 ones
        assert (bv.select(q1, 1) == want1).all() and bv.device_bytes == b1  # built once
        assert (bv.select(q0, 0) == want0).all() and bv.device_bytes > b1  # the zeros get their own
        bv.set_batch_order(pkg.ORDER_DIRECT)  # the direct kernel uses them too once they exist
        assert (bv.select(q1, 1) == want1).all() and (bv.select(q0, 0) == want0).all()
        every = np.arange(1, bv.arg_count(1) + 1, dtype=np.uint64)  # every one of the vector, sector boundaries included
        assert (bv.select(every, 1) == o.select(every, 1)).all()
        for other in (compact, sp):
            other.set_batch_order(pkg.ORDER_BINNED)
            before = other.device_bytes
            q = cases.select_queries(other.arg_count(1), 5, 20000)
            got = other.select(q, 1)
            assert other.device_bytes == before
            if other is compact:
                assert (got == o.select(q, 1)).all()
            else:
                assert (got == oracle.bv(sparse, nbits).select(q, 1)).all()
