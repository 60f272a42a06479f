"""The SDSL-shaped C++ header (sdsl-lite_b200/include/sdsl_b200.hpp): it must compile against the C ABI on a
CPU-only box, and its results must match naive scans on the GPU (tests/cpp/shim_test.cpp)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "sdsl-lite_b200")
BIN = os.path.join(PKG, "build", "shim_test")


def _build(pkg):
    if not os.path.exists(pkg.LIB_PATH):
        pkg.build()
    os.makedirs(os.path.dirname(BIN), exist_ok=True)
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", os.path.join(ROOT, "tests", "cpp", "shim_test.cpp"), "-o", BIN,
           "-L" + PKG, "-lsdslgpu", "-Wl,-rpath," + PKG]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]


def test_shim_compiles_and_links(pkg):
    _build(pkg)
    assert os.path.exists(BIN)


def test_shim_host_only_parts(pkg):
    """bit_vector storage and its serialised form need no device"""
    _build(pkg)
    r = subprocess.run([BIN, "--host-only"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "host-only ok" in r.stdout, (r.stdout[-2000:], r.stderr[-2000:])


@pytest.mark.gpu
def test_shim_matches_naive_on_gpu(pkg):
    _build(pkg)
    r = subprocess.run([BIN], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "shim_test ok" in r.stdout, (r.stdout[-2000:], r.stderr[-2000:])
