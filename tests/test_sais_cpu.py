"""CPU: sdsl-lite_b200/csrc/sais.h — the host SA-IS suffix sorter the FM-index builder falls back to when the text
does not fit 32-bit suffix indices or device memory — against a naive suffix sort on small texts and against the
suffix array inside the oracle's / the reference's index (csa[i], csa_wt.hpp:363-381) on the text catalogue."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import texts

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "sdsl-lite_b200", "build", "libsaishost.so")


@pytest.fixture(scope="module")
def sais():
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    src = os.path.join(ROOT, "tests", "cpp", "sais_host.cpp")
    hdr = os.path.join(ROOT, "sdsl-lite_b200", "csrc", "sais.h")
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        r = subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-shared", "-fPIC", src, "-o", SO], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-3000:]
    L = ctypes.CDLL(SO)
    L.sais_host.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int, ctypes.c_void_p]

    def run(t, wide=0):
        a = np.frombuffer(t, dtype=np.uint8)
        a = np.ascontiguousarray(a) if len(a) else np.zeros(1, np.uint8)
        out = np.zeros(len(t) + 1, np.uint64)
        L.sais_host(a.ctypes.data, len(t), wide, out.ctypes.data)
        return out

    return run


def test_sais_vs_naive(sais):
    rng = np.random.default_rng(3)
    cases_ = [b"", b"a", b"aa", b"ab", b"ba", b"banana", b"mississippi", b"abracadabra", b"a" * 50, b"ab" * 40, bytes(range(1, 256))]
    cases_ += [rng.integers(1, 1 + int(s), int(n), dtype=np.uint8).tobytes() for s, n in [(1, 30), (2, 200), (3, 500), (4, 999), (255, 700)]]
    for t in cases_:
        full = t + b"\x00"
        want = sorted(range(len(full)), key=lambda i: full[i:])
        for wide in (0, 1):
            assert list(sais(t, wide)) == want, (t[:20], wide)


def test_sais_vs_index_suffix_array(sais, oracle, orc):
    mk = orc.Ref() if orc.ref_available() else oracle
    for name, t in texts.text_catalogue(zero_free=True, large=False):
        sa = sais(t)
        idx = np.arange(len(t) + 1, dtype=np.uint64)
        assert (mk.csa(t).sa(idx) == sa).all(), name
        assert (sais(t, 1) == sa).all(), (name, "64-bit indices")
