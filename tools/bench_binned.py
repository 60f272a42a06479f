#!/usr/bin/env python
"""direct vs binned execution of the BASELINE config[1] batch (2^33-bit vector, 1e8 uniform queries), chunk-size sweep.
One JSON line per (op, order, chunk) with CUDA-event times; results compared bit for bit between the orders."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--nbits-log2", type=int, default=33)
ap.add_argument("--queries", type=float, default=1e8)
ap.add_argument("--chunks", default="8,16,32")
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--density", type=float, default=0.5)
ap.add_argument("--ops", default="rank1,select1,rank0")
ap.add_argument("--out", default=None)
args = ap.parse_args()

pkg = ge.load_package()
nbits, nq = 1 << args.nbits_log2, int(args.queries)
g = torch.Generator(device="cuda").manual_seed(42)
words = torch.randint(-(2**63), 2**63 - 1, (nbits // 64,), dtype=torch.int64, device="cuda", generator=g)
if args.density < 0.5:
    k = int(round(-np.log2(args.density)))
    for _ in range(k - 1):
        words &= torch.randint(-(2**63), 2**63 - 1, (nbits // 64,), dtype=torch.int64, device="cuda", generator=g)
bv = pkg.BitVector(words, nbits)
del words
idx = torch.randint(0, nbits + 1, (nq,), dtype=torch.int64, device="cuda", generator=g)
m = bv.arg_count(1)
sel = torch.randint(1, m + 1, (nq,), dtype=torch.int64, device="cuda", generator=g)
out = torch.empty(nq, dtype=torch.int64, device="cuda")
lines = []


def timed(fn):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(args.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), sorted(ts)[len(ts) // 2]


ref = {}
for op in args.ops.split(","):
    q = sel if op.startswith("select") else idx
    b = 0 if op.endswith("0") else 1
    call = (lambda: bv.rank(q, b, out=out)) if op.startswith("rank") else (lambda: bv.select(q, b, out=out))
    bv.set_batch_order(pkg.ORDER_DIRECT)
    best, med = timed(call)
    ref[op] = out.clone()
    lines.append({"op": op, "order": "direct", "ms_best": best, "ms_median": med, "gqps": nq / best / 1e6})
    print(json.dumps(lines[-1]), flush=True)
    bv.set_batch_order(pkg.ORDER_BINNED)
    for c in args.chunks.split(","):
        os.environ["SDSLGPU_BIN_CHUNK_BYTES"] = str(int(float(c) * (1 << 20)))
        best, med = timed(call)
        same = bool((out == ref[op]).all())
        lines.append({"op": op, "order": "binned", "chunk_mib": float(c), "ms_best": best, "ms_median": med, "gqps": nq / best / 1e6, "bit_exact_vs_direct": same})
        print(json.dumps(lines[-1]), flush=True)
    os.environ.pop("SDSLGPU_BIN_CHUNK_BYTES", None)
if args.out:
    with open(args.out, "w") as f:
        for ln in lines:
            f.write(json.dumps(ln) + "\n")
