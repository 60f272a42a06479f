#!/usr/bin/env python
"""select stride x interpolation on (a) a 2^33-bit random vector, (b) wt_huff.select on 2^28 uniform bytes"""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import __graft_entry__ as ge
from bench_all import dev, gpu_random_words, time_gpu
pkg = ge.load_package()
nbits, nq = 1 << 33, int(1e8)
words = gpu_random_words(nbits, 0.5, 5)
n = 1 << 28
text = np.random.default_rng(42).integers(0, 256, n, dtype=np.uint8)
occ = np.bincount(text, minlength=256).astype(np.uint64)
for interp in ("0", "1"):
    for ls in ("5", "6", "7", "8", "9", "10"):
        os.environ["SDSLGPU_SELECT_LOG_S"] = ls
        os.environ["SDSLGPU_SELECT_INTERP"] = interp
        bv = pkg.BitVector(words, nbits)
        sel = dev(np.random.default_rng(7).integers(1, bv.arg_count(1) + 1, nq, dtype=np.uint64))
        out = torch.empty(nq, dtype=torch.int64, device="cuda")
        ms, _ = time_gpu(lambda: bv.select(sel, 1, out=out), 5)
        bv.close()
        wt = pkg.WtHuff(text)
        nw = int(1e7)
        qr = np.random.default_rng(7)
        c_h = qr.integers(0, 256, nw, dtype=np.uint8)
        k = dev((qr.integers(0, 2**62, nw, dtype=np.uint64) % occ[c_h.astype(np.int64)]) + np.uint64(1))
        c = dev(c_h)
        out2 = torch.empty(nw, dtype=torch.int64, device="cuda")
        ms2, _ = time_gpu(lambda: wt.select(k, c, out=out2), 5)
        wt.close()
        print(json.dumps({"interp": int(interp), "log_s": int(ls), "bv_select_ms": ms, "bv_gqps": nq / ms / 1e6, "wt_select_ms": ms2, "wt_gqps": nw / ms2 / 1e6}), flush=True)
