"""CPU: the per-query DEVICE functions themselves — sdsl-lite_b200/csrc/bv_device.cuh (bv_rank1, bv_rank1_and_bit,
bv_bit, bv_select<B>: sampled hint, interpolated first probe, walk, bisection) and the word-level helpers of
common.cuh — compiled as plain C++ (SDSLGPU_HOST_EMU, tests/cpp/device_on_host.cpp) and checked against the oracle:
the same source the kernels run, every sample stride from 1 to 4096, interpolation on and off, random / sparse /
clustered data.  The GPU tests check the kernels; this checks their logic where no GPU is."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "sdsl-lite_b200", "build", "libdevhost.so")


@pytest.fixture(scope="module")
def emu():
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    src = os.path.join(ROOT, "tests", "cpp", "device_on_host.cpp")
    deps = [src, os.path.join(ROOT, "tests", "cpp", "host_image.h")] + [os.path.join(ROOT, "sdsl-lite_b200", "csrc", f) for f in ("bv_device.cuh", "common.cuh")]
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(d) for d in deps):
        r = subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-Wno-unknown-pragmas", "-DSDSLGPU_HOST_EMU", "-shared", "-fPIC", src, "-o", SO],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-3000:]
    L = ctypes.CDLL(SO)
    vp, u64, u32 = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint32
    L.emu_rank1.argtypes = [vp, u64, vp, u64, vp, vp]
    L.emu_select.argtypes = [vp, u64, ctypes.c_int, u32, u32, u32, vp, u64, vp]
    L.emu_select_sectors.argtypes = [vp, u64, ctypes.c_int, u32, vp, u64, vp]
    L.emu_select_sectors.restype = ctypes.c_int64
    L.emu_sect_div.argtypes = [u32, u64]
    L.emu_sect_div.restype = u64
    L.emu_sect_stride.argtypes = [u64, u64]
    L.emu_sect_stride.restype = u32
    L.emu_sel64.argtypes = [u64, u32]
    L.emu_sel64.restype = u32
    return L


def _words(w):
    return np.ascontiguousarray(np.concatenate([np.asarray(w, np.uint64), np.zeros(2, np.uint64)]))


def _shapes():
    for cid, w, nbits in cases.bitvector_catalogue(large=True):
        if nbits <= 1_100_000:
            yield cid, w, nbits
    rng = np.random.default_rng(77)
    n = 700_000
    b = np.zeros(n, np.uint8)
    b[-5000:] = 1  # every one at the very end: the interpolated guess is far too low, bisection territory
    yield "ones_at_end", cases.pack_bits(b), n
    b = np.zeros(n, np.uint8)
    b[:3000] = 1
    b[350_000:353_000] = 1
    b[-1] = 1  # two clusters and a straggler: both over- and undershoots over long hint ranges
    yield "clusters", cases.pack_bits(b), n
    b = (rng.random(n) < 0.5).astype(np.uint8)
    b[100_000:600_000] = 0  # a desert inside random data (and an oasis for the zeros)
    yield "desert", cases.pack_bits(b), n
    # gaps of 2 - 15 Mbit between neighbouring ones: hint ranges of tens of thousands of blocks around a handful of ones
    n = 36_000_000
    b = np.zeros(n, np.uint8)
    b[[0, 5, 2_200_000, 2_200_001, 4_500_000, 19_500_000, 19_500_003, 35_999_999]] = 1
    b[7_000_000:7_000_300] = rng.random(300) < 0.5
    yield "wide_gaps", cases.pack_bits(b), n
    yield "wide_gaps_inverted", cases.pack_bits(1 - b), n


def test_word_select(emu):
    rng = np.random.default_rng(5)
    for x in list(rng.integers(1, 2**64, 400, dtype=np.uint64)) + [np.uint64(1), np.uint64(1 << 63), np.uint64(2**64 - 1)]:
        x = int(x)
        pos = [i for i in range(64) if (x >> i) & 1]
        for k, p in enumerate(pos, 1):  # test/bits_test.cpp:150-177
            assert emu.emu_sel64(x, k) == p


def test_device_rank_and_bit_logic(emu, oracle):
    for cid, w, nbits in _shapes():
        ww = _words(w)
        idx = cases.rank_queries(nbits, 9, 20000)
        out = np.zeros(len(idx), np.uint64)
        bit = np.full(len(idx), 7, np.uint64)
        emu.emu_rank1(ww.ctypes.data, nbits, idx.ctypes.data, len(idx), out.ctypes.data, bit.ctypes.data)
        assert (out == oracle.bv(w, nbits).rank(idx, 1)).all(), cid
        inside = idx < nbits
        if inside.any():
            bits = cases.unpack_bits(w, nbits)
            assert (bit[inside] == bits[idx[inside].astype(np.int64)]).all(), (cid, "bit / rank_and_bit")


@pytest.mark.parametrize("interp", [0, 1, 2])  # 2: position-valued samples
@pytest.mark.parametrize("log_s", [0, 3, 6, 9, 12])
def test_device_select_logic(emu, oracle, log_s, interp):
    checked = 0
    for cid, w, nbits in _shapes():
        ww = _words(w)
        ob = oracle.bv(w, nbits)
        for b in (1, 0):
            m = int(ob.rank([nbits], b)[0])
            q = cases.select_queries(m, 11 + log_s, 6000)
            if not len(q):
                continue
            out = np.zeros(len(q), np.uint64)
            emu.emu_select(ww.ctypes.data, nbits, b, log_s, interp & 1, interp >> 1, q.ctypes.data, len(q), out.ctypes.data)
            assert (out == ob.select(q, b)).all(), (cid, b, log_s, interp)
            checked += len(q)
    assert checked > 100000


def test_select_sector_stride_and_division(emu):
    """The sector index of a query is key / stride by one multiply-high (bv_sect_magic): exact for every stride the
    library can pick and every key a 2^36-bit vector can produce; the stride itself follows the density."""
    rng = np.random.default_rng(8)
    keys = [0, 1, 2**32 - 1, 2**32, 2**36 - 1, 2**40] + [int(x) for x in rng.integers(0, 2**36, 300, dtype=np.uint64)]
    for stride in list(range(8, 209)) + [2, 3, 7, 255, 256, 1000]:
        edge = [stride * k + d for k in (1, 12345, 2**30 // stride) for d in (-1, 0, 1)]
        for key in keys + edge:
            assert emu.emu_sect_div(stride, key) == key // stride, (stride, key)
    n = 1 << 33
    assert emu.emu_sect_stride(n // 2, n) == 81 and emu.emu_sect_stride(n // 4, n) == 34  # the measured optimum at density 1/2
    assert emu.emu_sect_stride(n // 50, n) == 0 and emu.emu_sect_stride(0, n) == 0        # too sparse for sectors
    assert 150 <= emu.emu_sect_stride(n, n) <= 208                                         # all ones: as many as a sector holds
    strides = [emu.emu_sect_stride(int(n * d), n) for d in (0.09, 0.1, 0.2, 0.3, 0.5, 0.7, 0.9)]
    assert strides == sorted(strides) and strides[0] >= 8


@pytest.mark.parametrize("stride", [0, 8, 64, 81, 208])  # 0: the stride the library picks from the density
def test_device_select_sectors(emu, oracle, stride):
    """Select sectors (bv_device.cuh: bv_make_sector / bv_select_sector): one 32-byte record answers a query, sectors
    whose B-bits do not fit are marked and answered by the sampled select.  Forced strides put dense sectors on sparse
    data (everything marked) and sparse sectors on dense data (nothing marked)."""
    checked = marked_total = built = 0
    for cid, w, nbits in _shapes():
        ww = _words(w)
        ob = oracle.bv(w, nbits)
        for b in (1, 0):
            m = int(ob.rank([nbits], b)[0])
            q = cases.select_queries(m, 23 + stride, 6000)
            if not len(q):
                continue
            out = np.zeros(len(q), np.uint64)
            marked = emu.emu_select_sectors(ww.ctypes.data, nbits, b, stride, q.ctypes.data, len(q), out.ctypes.data)
            if marked < 0:
                assert stride == 0, (cid, b)  # only the library's own choice may decline (density below ~8 %)
                continue
            built += 1
            assert (out == ob.select(q, b)).all(), (cid, b, stride)
            checked += len(q) - marked
            marked_total += marked
    assert built >= 10 and checked > 50000
    if stride:
        assert marked_total > 0  # the fallback was exercised


# ---------------------------------------------------------------------------------------------------------------
# rrr_vector<63>: rrr_device.cuh (rank, rank + bit, select<0/1>, enumerative decode) over an image built by the
# product's own host code (rrr_records_host, host_tables) from the reference's / the oracle's serialised vector
# ---------------------------------------------------------------------------------------------------------------
RRR_SO = os.path.join(ROOT, "sdsl-lite_b200", "build", "librrrhost.so")


@pytest.fixture(scope="module")
def rrr_emu():
    src = os.path.join(ROOT, "tests", "cpp", "rrr_on_host.cpp")
    deps = [src] + [os.path.join(ROOT, "sdsl-lite_b200", "csrc", f) for f in ("rrr_device.cuh", "common.cuh")]
    os.makedirs(os.path.dirname(RRR_SO), exist_ok=True)
    if not os.path.exists(RRR_SO) or os.path.getmtime(RRR_SO) < max(os.path.getmtime(d) for d in deps):
        r = subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-Wno-unknown-pragmas", "-DSDSLGPU_HOST_EMU", "-shared", "-fPIC", src, "-o", RRR_SO],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-3000:]
    L = ctypes.CDLL(RRR_SO)
    vp, u64 = ctypes.c_void_p, ctypes.c_uint64
    L.rrr_emu_load.restype = vp
    L.rrr_emu_load.argtypes = [vp]
    L.rrr_emu_free.argtypes = [vp]
    L.rrr_emu_rank.argtypes = [vp, ctypes.c_int, vp, u64, vp]
    L.rrr_emu_select.argtypes = [vp, ctypes.c_int, vp, u64, vp]
    L.rrr_emu_access.argtypes = [vp, vp, u64, vp, vp]
    return L


def _rrr_shapes():
    for cid, w, nbits in _shapes():
        yield cid, w, nbits
    # the density sweep of BASELINE config 3 in miniature; > 0.5 exercises inverted superblocks; 0.16 / 0.84 put the classes
    # on both sides of the switch between the one-at-a-time search (few ones / few zeros) and the position walk
    for d in (0.01, 0.05, 0.1, 0.16, 0.5, 0.84, 0.9, 0.95, 0.99):
        n = 300_000 + int(d * 1000)
        yield f"bernoulli.{d}", cases.bernoulli_words(n, d, 600 + int(d * 100)), n


def test_device_rrr_logic(rrr_emu, oracle):
    checked = 0
    for cid, w, nbits in _rrr_shapes():
        o = oracle.rrr(w, nbits)
        blob = np.frombuffer(o.serialize(), dtype=np.uint8).copy()
        blob = np.concatenate([blob, np.zeros(64, np.uint8)])
        h = rrr_emu.rrr_emu_load(blob.ctypes.data)
        try:
            idx = cases.rank_queries(nbits, 13, 8000)
            out = np.zeros(len(idx), np.uint64)
            for b in (1, 0):
                rrr_emu.rrr_emu_rank(h, b, idx.ctypes.data, len(idx), out.ctypes.data)
                assert (out == o.rank(idx, b)).all(), (cid, "rank", b)
                m = int(o.rank([nbits], b)[0])
                q = cases.select_queries(m, 14, 8000)
                if len(q):
                    so = np.zeros(len(q), np.uint64)
                    rrr_emu.rrr_emu_select(h, b, q.ctypes.data, len(q), so.ctypes.data)
                    assert (so == o.select(q, b)).all(), (cid, "select", b)
                    checked += len(q)
            pos = np.ascontiguousarray(idx[idx < nbits])
            if len(pos):
                bit, rk = np.zeros(len(pos), np.uint64), np.zeros(len(pos), np.uint64)
                rrr_emu.rrr_emu_access(h, pos.ctypes.data, len(pos), bit.ctypes.data, rk.ctypes.data)
                assert (bit == o.access(pos)).all() and (rk == o.rank(pos, 1)).all(), (cid, "rank_and_bit")
        finally:
            rrr_emu.rrr_emu_free(h)
    assert checked > 100000


# ---------------------------------------------------------------------------------------------------------------
# wt_huff<>: wt_device.cuh (wt_rank_one, wt_inverse_select_one) over the tree shape and bit planes of the product's
# host code (wt_shape.h) and the sector-block rank of bv_device.cuh
# ---------------------------------------------------------------------------------------------------------------
WT_SO = os.path.join(ROOT, "sdsl-lite_b200", "build", "libwthost.so")


@pytest.fixture(scope="module")
def wt_emu():
    src = os.path.join(ROOT, "tests", "cpp", "wt_on_host.cpp")
    deps = [src, os.path.join(ROOT, "tests", "cpp", "host_image.h")] + [
        os.path.join(ROOT, "sdsl-lite_b200", "csrc", f) for f in ("wt_device.cuh", "wt_shape.h", "wt_tree.h", "bits_access.cuh", "bv_device.cuh", "rrr_device.cuh", "common.cuh")]
    os.makedirs(os.path.dirname(WT_SO), exist_ok=True)
    if not os.path.exists(WT_SO) or os.path.getmtime(WT_SO) < max(os.path.getmtime(d) for d in deps):
        r = subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-Wno-unknown-pragmas", "-pthread", "-DSDSLGPU_HOST_EMU", "-shared", "-fPIC", src, "-o", WT_SO],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-3000:]
    L = ctypes.CDLL(WT_SO)
    vp, u64 = ctypes.c_void_p, ctypes.c_uint64
    L.wt_emu_create.restype = vp
    L.wt_emu_create.argtypes = [vp, u64]
    L.wt_emu_free.argtypes = [vp]
    L.wt_emu_rank.argtypes = [vp, vp, vp, u64, vp]
    L.wt_emu_inverse_select.argtypes = [vp, vp, u64, vp, vp]
    return L


def test_device_wt_logic(wt_emu, oracle):
    import texts

    rng = np.random.default_rng(23)
    deep = np.concatenate([np.full(1 << k, 65 + k, np.uint8) for k in range(14)])
    rng.shuffle(deep)
    for name, t in list(texts.text_catalogue(large=False)) + [("deep", deep.tobytes())]:
        a = np.ascontiguousarray(np.frombuffer(t, dtype=np.uint8))
        h = wt_emu.wt_emu_create(a.ctypes.data, len(a))
        try:
            o = oracle.wt_huff(t)
            i, c = texts.wt_queries(t, rng, 6000)
            out = np.zeros(len(i), np.uint64)
            wt_emu.wt_emu_rank(h, i.ctypes.data, c.ctypes.data, len(i), out.ctypes.data)
            assert (out == o.rank(i, c)).all(), (name, "rank")
            j = rng.integers(0, len(t), 6000, dtype=np.uint64)
            rk, sym = np.zeros(len(j), np.uint64), np.zeros(len(j), np.uint64)
            wt_emu.wt_emu_inverse_select(h, j.ctypes.data, len(j), rk.ctypes.data, sym.ctypes.data)
            orr, oss = o.inverse_select(j)
            assert (rk == orr).all() and (sym == oss).all(), (name, "inverse_select")
        finally:
            wt_emu.wt_emu_free(h)


# ---------------------------------------------------------------------------------------------------------------
# sd_vector<>: sd_device.cuh (rank via select_0 on `high` + the backwards scan, select_1) over an image built from the
# serialised vector
# ---------------------------------------------------------------------------------------------------------------
SD_SO = os.path.join(ROOT, "sdsl-lite_b200", "build", "libsdhost.so")


@pytest.fixture(scope="module")
def sd_emu():
    src = os.path.join(ROOT, "tests", "cpp", "sd_on_host.cpp")
    deps = [src, os.path.join(ROOT, "tests", "cpp", "host_image.h")] + [os.path.join(ROOT, "sdsl-lite_b200", "csrc", f) for f in ("sd_device.cuh", "bv_device.cuh", "common.cuh")]
    os.makedirs(os.path.dirname(SD_SO), exist_ok=True)
    if not os.path.exists(SD_SO) or os.path.getmtime(SD_SO) < max(os.path.getmtime(d) for d in deps):
        r = subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-Wno-unknown-pragmas", "-DSDSLGPU_HOST_EMU", "-shared", "-fPIC", src, "-o", SD_SO],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-3000:]
    L = ctypes.CDLL(SD_SO)
    vp, u64 = ctypes.c_void_p, ctypes.c_uint64
    L.sd_emu_load.restype = vp
    L.sd_emu_load.argtypes = [vp]
    L.sd_emu_free.argtypes = [vp]
    L.sd_emu_rank.argtypes = [vp, ctypes.c_int, vp, u64, vp]
    L.sd_emu_select1.argtypes = [vp, vp, u64, vp]
    L.sd_emu_select0.argtypes = [vp, ctypes.c_int, vp, u64, vp, vp]
    return L


def test_device_sd_logic(sd_emu, oracle):
    checked = 0
    for cid, w, nbits in _rrr_shapes():
        if nbits == 0:
            continue
        o = oracle.sd(w, nbits)
        blob = np.concatenate([np.frombuffer(o.serialize(), dtype=np.uint8), np.zeros(64, np.uint8)])
        h = sd_emu.sd_emu_load(blob.ctypes.data)
        try:
            idx = cases.rank_queries(nbits, 17, 8000)
            out = np.zeros(len(idx), np.uint64)
            for b in (1, 0):
                sd_emu.sd_emu_rank(h, b, idx.ctypes.data, len(idx), out.ctypes.data)
                assert (out == o.rank(idx, b)).all(), (cid, "rank", b)
            m = int(o.rank([nbits], 1)[0])
            q = cases.select_queries(m, 18, 8000)
            if len(q):
                so = np.zeros(len(q), np.uint64)
                sd_emu.sd_emu_select1(h, q.ctypes.data, len(q), so.ctypes.data)
                assert (so == o.select(q, 1)).all(), (cid, "select_1")
                checked += len(q)
        finally:
            sd_emu.sd_emu_free(h)
    assert checked > 100000


@pytest.mark.parametrize("log_s", [-1, 0, 5, 12])
def test_device_sd_select0_samples(sd_emu, oracle, log_s):
    """sd_select0_one (sample table over the zeros of the vector -> crossing block of `high` -> bucket) against the
    oracle's select_support_sd<0>; every zero of the small shapes, random ones of the large; the walk limit may only
    be hit on clustered data, and then the binary-search fallback answers"""
    checked = fell = 0
    for cid, w, nbits in _rrr_shapes():
        if nbits == 0:
            continue
        o = oracle.sd(w, nbits)
        z = nbits - int(o.rank([nbits], 1)[0])
        if z == 0:
            continue
        blob = np.concatenate([np.frombuffer(o.serialize(), dtype=np.uint8), np.zeros(64, np.uint8)])
        h = sd_emu.sd_emu_load(blob.ctypes.data)
        try:
            q = np.arange(1, z + 1, dtype=np.uint64) if z <= 20000 else np.unique(np.concatenate(
                [cases.select_queries(z, 19, 20000), np.array([1, 2, z - 1, z], dtype=np.uint64)]))
            out = np.zeros(len(q), np.uint64)
            fb = ctypes.c_uint64(0)
            sd_emu.sd_emu_select0(h, log_s, q.ctypes.data, len(q), out.ctypes.data, ctypes.byref(fb))
            assert (out == o.select(q, 0)).all(), (cid, "select_0", log_s)
            checked += len(q)
            fell += fb.value
            if log_s == -1 and cid.startswith("rand"):
                assert fb.value == 0, (cid, "uniformly random data never needs the fallback")
        finally:
            sd_emu.sd_emu_free(h)
    assert checked > 100000 and fell < checked // 2
