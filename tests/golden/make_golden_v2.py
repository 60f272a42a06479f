#!/usr/bin/env python
"""tests/golden/make_golden_v2.py — generates tests/golden/golden_v2.npz with the UNMODIFIED reference
(oracle/_ref/libsdslref.so): SHA-256 digests of

  * rank_support_v5<1> / <0> tables (rank_support_v5.hpp:66-158) for the bit-vector catalogue,
  * the reference's count-benchmark index FM_HUFF (benchmark/indexing_count/index.config:8) =
    csa_wt<wt_huff<bit_vector, rank_support_v5<>, select_support_scan<>, select_support_scan<0>>, 1<<20, 1<<20>
    and of the wavelet tree inside it, for the zero-free texts, plus its count() answers for seeded patterns.

Run in the build container:  python tests/golden/make_golden_v2.py"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
import texts  # noqa: E402
from oracle.pyoracle import Ref, csr_patterns  # noqa: E402


def sha(b):
    return np.frombuffer(hashlib.sha256(b).digest(), dtype=np.uint8)


def main():
    r = Ref()
    out = {}
    for cid, w, nbits in cases.bitvector_catalogue(large=True):
        if nbits > 1_000_000:
            continue
        bv = r.bv(w, nbits, with_select=False)
        out[f"v5|{cid}|sha"] = np.concatenate([sha(bv.serialize(5)), sha(bv.serialize(6))])
    rng = np.random.default_rng(2025)
    for name, t in texts.text_catalogue(zero_free=True, large=False):
        blob, count = r.fm_huff(text=t)
        pats = [t[s : s + int(rng.integers(1, 12))] for s in rng.integers(0, max(1, len(t) - 12), 150)] + [b"", b"\x01\x02zz"]
        flat, off = csr_patterns(pats)
        key = f"fm_huff|{name}"
        out[key + "|flat"], out[key + "|off"], out[key + "|cnt"] = flat, off, count(flat, off)
        out[key + "|sha"] = sha(blob)
        out[key + "|wt_sha"] = sha(r.wt_huff_v5_blob(t))
    path = os.path.join(HERE, "golden_v2.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
