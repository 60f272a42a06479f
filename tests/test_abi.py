"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/sdslgpu.h declares.
No compute call is made here (there is no GPU in the build container)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "sdslgpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sdslgpu_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(pkg):
    if not os.path.exists(pkg.LIB_PATH):
        pkg.build()
    L = pkg.lib()
    declared = header_symbols()
    assert len(declared) >= 10
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/sdslgpu.h but not exported"
    # the python binding knows every declared symbol too (so a new entry point cannot go untested)
    assert sorted(pkg.declared_symbols()) == declared
    assert b"sm_100a" in L.sdslgpu_version()


def test_fails_loudly_without_gpu(pkg):
    """the product path must not fall back to the CPU: without a device, creation raises"""
    import ctypes as C

    n = C.c_int(-1)
    st = pkg.lib().sdslgpu_device_count(C.byref(n))
    if st == 0 and n.value > 0:
        pytest.skip("a CUDA device is present")
    import numpy as np

    with pytest.raises(pkg.SdslGpuError):
        pkg.BitVector(np.zeros(4, np.uint64), 200)


def test_sass_is_sm100a_with_sector_loads(pkg):
    """the rank kernel is real sm_100a SASS with one 256-bit gather per query"""
    import shutil
    import subprocess

    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not on PATH")
    out = subprocess.run(["cuobjdump", "-sass", pkg.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert re.search(r"LDG\.E\.[A-Z.0-9]*256", out), "no 256-bit LDG in the SASS"
