#!/usr/bin/env python
"""tests/golden/make_golden.py — generates tests/golden/golden_v1.npz with the UNMODIFIED reference
(oracle/_ref/libsdslref.so, built from /root/reference/include by oracle/Makefile).  Run in the build container:

    python tests/golden/make_golden.py

The fixture lets the oracle (and the CUDA path) be checked against the reference's answers on machines where the
reference library is absent.  Inputs are regenerated from tests/cases.py / tests/texts.py seeds; only queries,
answers and SHA-256 digests of the reference's serialised bytes are stored."""
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
from test_oracle_wt_int import queries as int_queries  # noqa: E402
from test_oracle_wt_int import sequences  # noqa: E402


def sha(b):
    return np.frombuffer(hashlib.sha256(b).digest(), dtype=np.uint8)


def golden_patterns(t, rng, k):
    n = len(t)
    pats = []
    for _ in range(k):
        m = int(rng.integers(1, 16))
        if n and rng.random() < 0.7:
            s = int(rng.integers(0, max(1, n - m + 1)))
            pats.append(t[s : s + m])
        else:
            pats.append(rng.integers(1, 256, m, dtype=np.uint8).tobytes())
    return pats + [b"", t[:3]]


def main():
    r = Ref()
    out = {}
    for cid, w, nbits in cases.bitvector_catalogue(large=True):
        if nbits > 1_000_000:
            continue
        idx = cases.rank_queries(nbits, 101, 300)
        for kind, obj in (("bv", r.bv(w, nbits)), ("rrr", r.rrr(w, nbits)), ("sd", r.sd(w, nbits) if nbits else None)):
            if obj is None:
                continue
            key = f"{kind}|{cid}"
            out[key + "|idx"] = idx
            for b in (0, 1):
                out[key + f"|rank{b}"] = obj.rank(idx, b)
                m = int(obj.rank([nbits], b)[0])
                q = cases.select_queries(m, 102, 300)
                out[key + f"|sel{b}_q"] = q
                out[key + f"|sel{b}"] = obj.select(q, b) if len(q) else q
            blobs = [obj.serialize(k) for k in range(5)] if kind == "bv" else [obj.serialize()]
            out[key + "|sha"] = np.concatenate([sha(x) for x in blobs])
    rng = np.random.default_rng(2024)
    for name, t in texts.text_catalogue(large=False):
        wt = r.wt_huff(t)
        i, c = texts.wt_queries(t, rng, 400)
        j = rng.integers(0, len(t), 400, dtype=np.uint64)
        rr, ss = wt.inverse_select(j)
        key = f"wt_huff|{name}"
        out[key + "|i"], out[key + "|c"], out[key + "|rank"] = i, c, wt.rank(i, c)
        out[key + "|j"], out[key + "|inv_rank"], out[key + "|inv_sym"] = j, rr, ss
        out[key + "|sel"] = wt.select(rr + np.uint64(1), ss.astype(np.uint8))
        out[key + "|sha"] = sha(wt.serialize())
    for name, seq in sequences():
        wt = r.wt_int(seq)
        i, c = int_queries(seq, rng, 400)
        j = rng.integers(0, len(seq), 400, dtype=np.uint64)
        rr, ss = wt.inverse_select(j)
        key = f"wt_int|{name}"
        out[key + "|i"], out[key + "|c"], out[key + "|rank"] = i, c, wt.rank(i, c)
        out[key + "|j"], out[key + "|inv_rank"], out[key + "|inv_sym"] = j, rr, ss
        out[key + "|sha"] = sha(wt.serialize())
    for name, t in texts.text_catalogue(zero_free=True, large=False):
        csa = r.csa(t)
        flat, off = csr_patterns(golden_patterns(t, rng, 120))
        cnt, l = csa.count(flat, off, want_l=True)
        keep = cnt <= 300
        pats = [flat[int(off[k]) : int(off[k + 1])].tobytes() for k in range(len(cnt)) if keep[k]]
        flat, off = csr_patterns(pats)
        cnt, l = csa.count(flat, off, want_l=True)
        occ_off, occ = csa.locate(flat, off)
        k = rng.integers(0, len(t) + 1, 300, dtype=np.uint64)
        key = f"csa|{name}"
        out[key + "|flat"], out[key + "|off"], out[key + "|cnt"], out[key + "|l"] = flat, off, cnt, l
        out[key + "|occ_off"], out[key + "|occ"] = occ_off, occ
        out[key + "|sa_i"], out[key + "|sa"] = k, csa.sa(k)
        out[key + "|sha"] = sha(csa.serialize())
    path = os.path.join(HERE, "golden_v1.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
