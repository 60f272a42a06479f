#!/usr/bin/env python
"""Where does the locality-ordered (binned) pipeline start to pay?  Batch-size sweep of rank_1 / select_1 on one
bit vector in both batch orders (sdslgpu_set_batch_order), CUDA-event times, results compared bit for bit between the
orders.  One JSON line per (op, n, order).  The AUTO threshold of csrc/binned.cu (bin_wanted) is set from this table.

    python tools/sweep_order.py --nbits-log2 33 --log2n 21,22,23,24,25,26 --out gpurun_out/sweep.jsonl
    SDSLGPU_SELECT_SECTORS=0 python tools/sweep_order.py ...         (select through the samples only)
"""
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
ap.add_argument("--log2n", default="21,22,23,24,25,26")
ap.add_argument("--extra-n", default="12500000,25000000,50000000,100000000")
ap.add_argument("--reps", type=int, default=7)
ap.add_argument("--ops", default="rank1,select1")
ap.add_argument("--tag", default="")
ap.add_argument("--out", default=None)
args = ap.parse_args()

pkg = ge.load_package()
nbits = 1 << args.nbits_log2
rng = np.random.default_rng(42)
words = rng.integers(0, 2**64, (nbits + 63) // 64, dtype=np.uint64)
bv = pkg.BitVector(words, nbits)
del words
m = bv.arg_count(1)
ns = sorted({1 << int(x) for x in args.log2n.split(",") if x} | {int(x) for x in args.extra_n.split(",") if x})
nmax = max(ns)
g = torch.Generator(device="cuda").manual_seed(7)
idx = torch.randint(0, nbits + 1, (nmax,), dtype=torch.int64, device="cuda", generator=g)
sel = torch.randint(1, m + 1, (nmax,), dtype=torch.int64, device="cuda", generator=g)
out = torch.empty(nmax, dtype=torch.int64, device="cuda")
ref = torch.empty(nmax, dtype=torch.int64, device="cuda")
sink = open(args.out, "a") if args.out else None


def timed(fn):
    for _ in range(2):
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
    ts.sort()
    return ts[0], ts[len(ts) // 2]


for op in args.ops.split(","):
    q_all = sel if op.startswith("select") else idx
    b = 0 if op.endswith("0") else 1
    for n in ns:
        q = q_all[:n]
        row = {"op": op, "n": n, "nbits_log2": args.nbits_log2, "tag": args.tag, "pos_samples": os.environ.get("SDSLGPU_SELECT_POS_SAMPLES", "default"),
               "select_sectors": os.environ.get("SDSLGPU_SELECT_SECTORS", "default")}
        for order, name in ((pkg.ORDER_DIRECT, "direct"), (pkg.ORDER_BINNED, "binned")):
            bv.set_batch_order(order)
            o = ref if name == "direct" else out
            call = (lambda: bv.rank(q, b, out=o[:n])) if op.startswith("rank") else (lambda: bv.select(q, b, out=o[:n]))
            best, med = timed(call)
            row[name + "_ms"] = best
            row[name + "_ms_median"] = med
            row[name + "_gqps"] = n / best / 1e6
        row["bit_exact_between_orders"] = bool((out[:n] == ref[:n]).all())
        row["binned_over_direct"] = row["direct_ms"] / row["binned_ms"]
        print(json.dumps(row), flush=True)
        if sink:
            sink.write(json.dumps(row) + "\n")
            sink.flush()
