#!/usr/bin/env python
"""Golden output of tests/cpp/fm_index_dropin.cpp built against the UNMODIFIED reference headers (run in the build
container, where /root/reference exists): writes fm_index_dropin.text / .queries / .expected next to this script.
tests/test_dropin.py builds the SAME source against sdsl-lite_b200/include/sdsl_b200.hpp and compares on the GPU."""
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import texts  # noqa: E402

REF_INC = os.environ.get("SDSL_REFERENCE_INCLUDE", "/root/reference/include")
QUERIES = ["ACGT", "GATTACA", "CTACGACCAGG", "TTTTTTTTTTTTTTTT", "NNN", "A", "", "CGTACG", "ACGTN", "GGGGGGGG"]


def main():
    text = bytes(dict(texts.text_catalogue(zero_free=True, large=False))["dna"])
    open(os.path.join(HERE, "fm_index_dropin.text"), "wb").write(text)
    open(os.path.join(HERE, "fm_index_dropin.queries"), "w").write("".join(q + "\n" for q in QUERIES))
    with tempfile.TemporaryDirectory() as tmp:
        exe = os.path.join(tmp, "ref_build")
        subprocess.run(["g++", "-std=c++17", "-O1", "-DUSE_REFERENCE", "-I" + REF_INC, os.path.join(ROOT, "tests", "cpp", "fm_index_dropin.cpp"), "-o", exe], check=True)
        out = subprocess.run([exe, os.path.join(HERE, "fm_index_dropin.text"), os.path.join(tmp, "idx")], stdin=open(os.path.join(HERE, "fm_index_dropin.queries")),
                             capture_output=True, check=True).stdout
    open(os.path.join(HERE, "fm_index_dropin.expected"), "wb").write(out)
    print(f"{len(out)} bytes of expected output")


if __name__ == "__main__":
    main()
