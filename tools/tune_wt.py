#!/usr/bin/env python
"""tools/tune_wt.py — wt_huff rank(i,c): per-query kernel vs the level-synchronous form (SDSLGPU_WT_LEVEL_SYNC),
on the BASELINE config-4 tree (2^28 uniform bytes) and a larger one, for several batch sizes."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import __graft_entry__ as ge  # noqa: E402
from bench_all import dev, time_gpu  # noqa: E402

pkg = ge.load_package()
for text_log2 in (28, 30):
    n = 1 << text_log2
    text = np.random.default_rng(42).integers(0, 256, n, dtype=np.uint8)
    wt = pkg.WtHuff(text)
    for nq in (int(1e6), int(1e7), int(1e8)):
        qr = np.random.default_rng(7)
        i, c = dev(qr.integers(0, n + 1, nq, dtype=np.uint64)), dev(qr.integers(0, 256, nq, dtype=np.uint8))
        res = {}
        for mode in ("0", "1"):
            os.environ["SDSLGPU_WT_LEVEL_SYNC"] = mode
            out = torch.empty(nq, dtype=torch.int64, device="cuda")
            ms, best = time_gpu(lambda: wt.rank(i, c, out=out), 5)
            res[mode] = out
            print(json.dumps({"text_log2": text_log2, "queries": nq, "level_sync": int(mode), "ms": ms, "gqps": nq / ms / 1e6,
                              "frac_of_hbm_peak": nq / ms / 1e6 * 209 / 6551.4}), flush=True)
        assert bool((res["0"] == res["1"]).all().item())
    wt.close()

# wt.select with the select-sample stride forced to the round-1a value (2^6) vs the automatic rule
n = 1 << 28
text = np.random.default_rng(42).integers(0, 256, n, dtype=np.uint8)
for ls in ("auto", "6", "8"):
    if ls == "auto":
        os.environ.pop("SDSLGPU_SELECT_LOG_S", None)
    else:
        os.environ["SDSLGPU_SELECT_LOG_S"] = ls
    wt = pkg.WtHuff(text)
    nq = int(1e7)
    qr = np.random.default_rng(7)
    c_h = qr.integers(0, 256, nq, dtype=np.uint8)
    occ = np.bincount(text, minlength=256).astype(np.uint64)
    k = dev((qr.integers(0, 2**62, nq, dtype=np.uint64) % occ[c_h.astype(np.int64)]) + np.uint64(1))
    c = dev(c_h)
    out = torch.empty(nq, dtype=torch.int64, device="cuda")
    ms, best = time_gpu(lambda: wt.select(k, c, out=out), 5)
    print(json.dumps({"op": "wt.select", "select_log_s": ls, "ms": ms, "gqps": nq / ms / 1e6}), flush=True)
    wt.close()
