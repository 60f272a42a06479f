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
ap.add_argument("--numpy-data", action="store_true", help="vector and queries generated exactly as bench.py does (numpy, host upload)")
ap.add_argument("--numpy-words", action="store_true")
ap.add_argument("--numpy-queries", action="store_true")
ap.add_argument("--alternate", action="store_true", help="time rank and select alternately in one stream like a bench.py step")
args = ap.parse_args()

pkg = ge.load_package()
nbits, nq = 1 << args.nbits_log2, int(args.queries)
g = torch.Generator(device="cuda").manual_seed(42)
args.numpy_words |= args.numpy_data
args.numpy_queries |= args.numpy_data
if True:
    rng = np.random.default_rng(42)
    words_np = rng.integers(0, 2**64, (nbits + 63) // 64, dtype=np.uint64)
    qr = np.random.default_rng(7)
    idx_np = qr.integers(0, nbits + 1, nq, dtype=np.uint64)
words = torch.from_numpy(words_np.view(np.int64)).cuda() if args.numpy_words else torch.randint(-(2**63), 2**63 - 1, (nbits // 64,), dtype=torch.int64, device="cuda", generator=g)
if args.density < 0.5:
    k = int(round(-np.log2(args.density)))
    for _ in range(k - 1):
        words &= torch.randint(-(2**63), 2**63 - 1, (nbits // 64,), dtype=torch.int64, device="cuda", generator=g)
bv = pkg.BitVector(words, nbits)
del words
idx = torch.randint(0, nbits + 1, (nq,), dtype=torch.int64, device="cuda", generator=g)
m = bv.arg_count(1)
sel = torch.randint(1, m + 1, (nq,), dtype=torch.int64, device="cuda", generator=g)
print(json.dumps({"ones": m, "numpy_words": args.numpy_words, "numpy_queries": args.numpy_queries}), flush=True)
if args.numpy_queries:
    idx = torch.from_numpy(idx_np.view(np.int64)).cuda()
    sel = torch.from_numpy(qr.integers(1, m + 1, nq, dtype=np.uint64).view(np.int64)).cuda()
if args.alternate:
    out2 = torch.empty(nq, dtype=torch.int64, device="cuda")
    out = torch.empty(nq, dtype=torch.int64, device="cuda")
    bv.set_batch_order(pkg.ORDER_BINNED)
    for _ in range(3):
        bv.rank(idx, 1, out=out)
        bv.select(sel, 1, out=out2)
    torch.cuda.synchronize()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(10)]
    for e in ev:
        e[0].record(); bv.rank(idx, 1, out=out); e[1].record(); bv.select(sel, 1, out=out2); e[2].record()
    torch.cuda.synchronize()
    print(json.dumps({"alternate": True, "numpy_data": args.numpy_data, "rank_ms": sum(e[0].elapsed_time(e[1]) for e in ev) / 10,
                      "select_ms": sum(e[1].elapsed_time(e[2]) for e in ev) / 10}), flush=True)
    sys.exit(0)
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
