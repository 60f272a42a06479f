"""CPU: the host-side writer of select_support_mcl blobs (sdsl-lite_b200/csrc/sdsl_pack.h, used by the library's
egress path sdslgpu_serialize) against the UNMODIFIED reference's serialize() — select_support_mcl.hpp:474-518
with the contents init_slow (:207-266) / init_fast (:269-381) produce.  The argument positions come from a naive
scan here (tests/cpp/pack_host.cpp) and from the batched select kernel on the GPU (tests/test_egress_gpu.py)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "sdsl-lite_b200", "build", "libpackhost.so")


@pytest.fixture(scope="module")
def packer():
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    src = os.path.join(ROOT, "tests", "cpp", "pack_host.cpp")
    hdr = os.path.join(ROOT, "sdsl-lite_b200", "csrc", "sdsl_pack.h")
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        r = subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-shared", "-fPIC", src, "-o", SO], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-3000:]
    L = ctypes.CDLL(SO)
    L.pack_select_mcl_naive.restype = ctypes.c_uint64
    L.pack_select_mcl_naive.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int, ctypes.c_void_p, ctypes.c_uint64]

    def pack(words, nbits, b):
        w = np.ascontiguousarray(np.concatenate([np.asarray(words, np.uint64), np.zeros(1, np.uint64)]))
        cap = 1 << 16
        while True:
            buf = np.zeros(cap, np.uint8)
            n = L.pack_select_mcl_naive(w.ctypes.data, nbits, b, buf.ctypes.data, cap)
            if n <= cap:
                return buf[:n].tobytes()
            cap = int(n)

    return pack


def boundary_cases():
    """argument counts around the rules' edges: 4032 / 4033 arguments in the trailing block (init_fast closes a block at
    its 4033rd argument), exact multiples of 4096, both sides of the 100000-bit init_slow / init_fast switch, spans
    around log^4 n"""
    rng = np.random.default_rng(12)
    for n in (99999, 100000, 300000):
        for m in (1, 63, 64, 65, 4031, 4032, 4033, 4034, 4095, 4096, 4097, 8191, 8192, 8193, 4096 + 4032, 4096 + 4033, 3 * 4096):
            bits = np.zeros(n, np.uint8)
            bits[rng.choice(n, m, replace=False)] = 1
            yield f"edge.{n}.{m}", cases.pack_bits(bits), n
    # dense prefix + sparse tail: mini and long blocks in one vector, in both modes
    for n in (90000, 2000000):
        bits = (rng.random(n) < 0.5).astype(np.uint8)
        bits[n // 3 :] = rng.random(n - n // 3) < 0.001
        yield f"mixed.{n}", cases.pack_bits(bits), n


@pytest.mark.parametrize("large", [False, True])
def test_select_mcl_blob_equals_reference(packer, ref, large):
    checked = 0
    for cid, w, nbits in cases.bitvector_catalogue(large=large):
        if large and nbits < 100000:
            continue
        if nbits > (1 << 22):
            continue  # the naive scan in the harness is slow; the big shapes run on the GPU
        if nbits % 64:
            w = w.copy()
            w[-1] &= np.uint64((1 << (nbits % 64)) - 1)  # the library masks the unspecified tail bits
        rb = ref.bv(w, nbits)
        for b in (1, 0):
            assert packer(w, nbits, b) == rb.serialize(3 if b else 4), (cid, b)
            checked += 1
    assert checked > 10


def test_select_mcl_blob_edges(packer, ref):
    for cid, w, nbits in boundary_cases():
        rb = ref.bv(w, nbits)
        for b in (1, 0):
            assert packer(w, nbits, b) == rb.serialize(3 if b else 4), (cid, b)


def test_select_mcl_blob_random_shapes(packer, ref):
    """seeded random shapes: sizes on both sides of the 100000-bit switch, densities from 1e-4 to 0.999, clustered runs
    (long superblocks next to mini ones), so that every combination of block kinds and widths shows up"""
    rng = np.random.default_rng(2026)
    kinds = set()
    for trial in range(48):
        n = int(rng.choice([rng.integers(1, 5000), rng.integers(60000, 100000), rng.integers(100000, 140000), rng.integers(400000, 1500000)]))
        d = float(rng.choice([1e-4, 1e-3, 0.01, 0.2, 0.5, 0.9, 0.999]))
        bits = (rng.random(n) < d).astype(np.uint8)
        if trial % 3 == 0:  # clusters: stretches of one value
            for _ in range(int(rng.integers(1, 6))):
                a = int(rng.integers(0, n))
                b = min(n, a + int(rng.integers(1, max(2, n // 3))))
                bits[a:b] = rng.integers(0, 2)
        w = cases.pack_bits(bits)
        rb = ref.bv(w, n)
        for b in (1, 0):
            blob = packer(w, n, b)
            assert blob == rb.serialize(3 if b else 4), (trial, n, d, b)
            m = int(np.frombuffer(blob[:8], np.uint64)[0])
            if m:  # which block kinds did this shape exercise? (mini_or_long is empty when no long block exists)
                sb_bits = int(np.frombuffer(blob[8:16], np.uint64)[0]) & ((1 << 56) - 1)
                pos = 16 + ((sb_bits + 63) // 64) * 8
                mol_bits = int(np.frombuffer(blob[pos : pos + 8], np.uint64)[0]) & ((1 << 56) - 1)
                kinds.add(("fast" if n >= 100000 else "slow", "mixed" if mol_bits else "mini-only"))
    assert len(kinds) == 4, kinds
