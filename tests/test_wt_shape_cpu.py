"""CPU: the host side of the wt_huff<> constructor (sdsl-lite_b200/csrc/wt_shape.h: Huffman shape with the reference's
tie-breaking, BFS numbering, per-symbol paths; multi-threaded fill of the bit planes) and the byte_tree writer of
sdsl_pack.h, against the UNMODIFIED reference's serialised tree (wt_pc.hpp:713-726): same sigma, same m_bv words, same
tree bytes.  On the GPU the same shape code feeds the device fill (wt_build.cu)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import texts

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "sdsl-lite_b200", "build", "libwtshapehost.so")


@pytest.fixture(scope="module")
def shape():
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    src = os.path.join(ROOT, "tests", "cpp", "wt_shape_host.cpp")
    deps = [src] + [os.path.join(ROOT, "sdsl-lite_b200", "csrc", f) for f in ("wt_shape.h", "wt_tree.h", "sdsl_pack.h")]
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(d) for d in deps):
        r = subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-pthread", "-shared", "-fPIC", src, "-o", SO], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-3000:]
    L = ctypes.CDLL(SO)
    vp, u64 = ctypes.c_void_p, ctypes.c_uint64
    L.wt_shape_host.restype = u64
    L.wt_shape_host.argtypes = [vp, u64, vp, u64, vp, u64, vp, vp]
    L.wt_shape_from_histogram.restype = u64
    L.wt_shape_from_histogram.argtypes = [vp, vp]

    def run(t):
        a = np.ascontiguousarray(np.frombuffer(t, dtype=np.uint8))
        cap_words = len(t) + 4  # a code has at most 56 bits (wt_helper.hpp: path length in 8 bits, path in 56)
        bv = np.zeros(cap_words, np.uint64)
        tree = np.zeros(8 + 511 * 22 + 512 + 2048, np.uint8)
        tb, sg = u64(), u64()
        bits = L.wt_shape_host(a.ctypes.data, len(t), bv.ctypes.data, cap_words, tree.ctypes.data, len(tree), ctypes.byref(tb), ctypes.byref(sg))
        assert (bits + 63) // 64 <= cap_words
        return int(bits), bv[: (bits + 63) // 64].tobytes(), tree[: tb.value].tobytes(), int(sg.value)

    run.lib = L
    return run


def _deep(rng):
    t = np.concatenate([np.full(1 << k, 65 + k, np.uint8) for k in range(16)])
    rng.shuffle(t)
    return t.tobytes()


def test_shape_bits_and_tree_equal_the_reference(shape, ref):
    rng = np.random.default_rng(41)
    cases_ = list(texts.text_catalogue(large=False)) + [("deep", _deep(rng)), ("ties", bytes(range(256)) * 3),
                                                        ("two_rare", b"a" * 5000 + b"b" + b"c")]
    for name, t in cases_:
        blob = ref.wt_huff(t).serialize()
        size, sigma = (int(x) for x in np.frombuffer(blob[:16], np.uint64))
        assert size == len(t)
        bits, bv, tree, sg = shape(t)
        assert sg == sigma, name
        hdr = int(np.frombuffer(blob[16:24], np.uint64)[0])
        assert hdr == (1 << 56) | bits, (name, "m_bv size")
        assert blob[24 : 24 + len(bv)] == bv, (name, "m_bv words")
        assert blob[-len(tree) :] == tree, (name, "byte_tree")


def test_code_depth_guard(shape):
    """Fibonacci weights give the deepest Huffman tree: k symbols -> depth k - 1.  57 symbols (depth 56) are the deepest
    the path words can hold; 58 must be refused, as the reference does (wt_helper.hpp:304-307)"""
    def depth_of(k):
        C = np.zeros(256, np.uint64)
        a, b = 1, 1
        for j in range(k):
            C[j] = a
            a, b = b, a + b
        d = ctypes.c_uint32(0)
        bits = shape.lib.wt_shape_from_histogram(C.ctypes.data, ctypes.byref(d))
        return bits, d.value

    bits, d = depth_of(57)
    assert bits != 2**64 - 1 and d == 56
    bits, d = depth_of(58)
    assert bits == 2**64 - 1
    bits, d = depth_of(20)
    assert d == 19
