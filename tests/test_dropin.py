"""Type-level drop-in of the C++ shim: tests/cpp/fm_index_dropin.cpp is ONE source that builds against the reference's
headers (-DUSE_REFERENCE) and against sdsl-lite_b200/include/sdsl_b200.hpp with only the include and the namespace
alias changed — `sdsl::csa_wt<sdsl::wt_huff<>> fm; construct_im(fm, text, 1); count / locate / extract / lex_interval /
backward_search; fm.C, fm.char2comp, fm.bwt.rank, fm.lf, fm.wavelet_tree; store_to_file / load_from_file`.
The reference build's output is the committed golden file (tests/golden/make_dropin_golden.py); the shim build must
print the same bytes on the GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "sdsl-lite_b200")
SRC = os.path.join(ROOT, "tests", "cpp", "fm_index_dropin.cpp")
GOLD = os.path.join(ROOT, "tests", "golden")
BIN = os.path.join(PKG, "build", "fm_index_dropin")
REF_INC = "/root/reference/include"


def _build_shim(pkg):
    if not os.path.exists(pkg.LIB_PATH):
        pkg.build()
    os.makedirs(os.path.dirname(BIN), exist_ok=True)
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", SRC, "-o", BIN, "-L" + PKG, "-lsdslgpu", "-Wl,-rpath," + PKG], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]


def test_same_source_compiles_against_the_shim(pkg):
    _build_shim(pkg)


def test_golden_is_what_the_reference_prints(tmp_path):
    """pins tests/golden/fm_index_dropin.expected to the unmodified reference (only where its headers exist)"""
    if not os.path.isdir(REF_INC):
        pytest.skip("/root/reference is not on this box")
    exe = str(tmp_path / "ref_build")
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-DUSE_REFERENCE", "-I" + REF_INC, SRC, "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]
    out = subprocess.run([exe, os.path.join(GOLD, "fm_index_dropin.text"), str(tmp_path / "idx")], stdin=open(os.path.join(GOLD, "fm_index_dropin.queries")),
                         capture_output=True, timeout=300).stdout
    assert out == open(os.path.join(GOLD, "fm_index_dropin.expected"), "rb").read()


@pytest.mark.gpu
def test_shim_build_prints_what_the_reference_prints(pkg, tmp_path):
    _build_shim(pkg)
    r = subprocess.run([BIN, os.path.join(GOLD, "fm_index_dropin.text"), str(tmp_path / "idx")], stdin=open(os.path.join(GOLD, "fm_index_dropin.queries")),
                       capture_output=True, timeout=600)
    want = open(os.path.join(GOLD, "fm_index_dropin.expected"), "rb").read()
    assert r.returncode == 0, r.stderr[-3000:]
    if r.stdout != want:
        got, exp = r.stdout.decode(errors="replace").splitlines(), want.decode().splitlines()
        diff = [(k, a, b) for k, (a, b) in enumerate(zip(got, exp)) if a != b][:5]
        raise AssertionError(f"{len(got)} vs {len(exp)} lines; first differences: {diff}")
