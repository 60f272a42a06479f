#!/usr/bin/env python
"""tools/tune_select.py — select_1 / select_0 rate on a 2^33-bit vector as a function of the sample stride
(SDSLGPU_SELECT_LOG_S) and density; evidence for the stride rule in csrc/bv.cu."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import __graft_entry__ as ge  # noqa: E402
from bench_all import dev, gpu_random_words, time_gpu  # noqa: E402

pkg = ge.load_package()
nbits, nq = 1 << 33, int(1e8)
for d in (0.5, 0.1, 0.01):
    words = gpu_random_words(nbits, d, 5)
    for ls in ("auto", 6, 8, 9, 10, 11, 12):
        if ls == "auto":
            os.environ.pop("SDSLGPU_SELECT_LOG_S", None)
        else:
            os.environ["SDSLGPU_SELECT_LOG_S"] = str(ls)
        bv = pkg.BitVector(words, nbits)
        for b in (1, 0):
            m = bv.arg_count(b)
            sel = dev(np.random.default_rng(7).integers(1, m + 1, nq, dtype=np.uint64))
            out = torch.empty(nq, dtype=torch.int64, device="cuda")
            ms, best = time_gpu(lambda: bv.select(sel, b, out=out), 5)
            chk = bool((bv.rank(out, b) == sel - 1).all().item())
            print(json.dumps({"density": d, "log_s": ls, "b": b, "ms": ms, "gqps": nq / ms / 1e6, "rank(select(k))==k-1": chk,
                              "index_MiB": bv.device_bytes / 2**20}), flush=True)
        bv.close()
