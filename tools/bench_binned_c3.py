#!/usr/bin/env python
"""direct vs binned rank / select_1 on rrr_vector<63> and sd_vector<> (BASELINE config 3 shape: 2^33 bits, density sweep)."""
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
ap.add_argument("--densities", default="0.5,0.125")
ap.add_argument("--kinds", default="sd,rrr")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--out", default=None)
args = ap.parse_args()
pkg = ge.load_package()
nbits, nq = 1 << args.nbits_log2, int(args.queries)
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
    return min(ts)


for d in [float(x) for x in args.densities.split(",")]:
    g = torch.Generator(device="cuda").manual_seed(42)
    words = torch.randint(-(2**63), 2**63 - 1, (nbits // 64,), dtype=torch.int64, device="cuda", generator=g)
    for _ in range(int(round(-np.log2(d))) - 1):
        words &= torch.randint(-(2**63), 2**63 - 1, (nbits // 64,), dtype=torch.int64, device="cuda", generator=g)
    idx = torch.randint(0, nbits + 1, (nq,), dtype=torch.int64, device="cuda", generator=g)
    out = torch.empty(nq, dtype=torch.int64, device="cuda")
    for kind in args.kinds.split(","):
        v = (pkg.SdVector if kind == "sd" else pkg.RrrVector)(words, nbits)
        m = v.arg_count(1)
        sel = torch.randint(1, m + 1, (nq,), dtype=torch.int64, device="cuda", generator=g)
        for op, q in (("rank1", idx), ("select1", sel)):
            call = (lambda: v.rank(q, 1, out=out)) if op == "rank1" else (lambda: v.select(q, 1, out=out))
            v.set_batch_order(pkg.ORDER_DIRECT)
            t_d = timed(call)
            ref = out.clone()
            v.set_batch_order(pkg.ORDER_BINNED)
            t_b = timed(call)
            same = bool((out == ref).all())
            lines.append({"kind": kind, "density": d, "op": op, "direct_ms": t_d, "binned_ms": t_b, "direct_gqps": nq / t_d / 1e6,
                          "binned_gqps": nq / t_b / 1e6, "bit_exact": same, "index_bytes": v.device_bytes})
            print(json.dumps(lines[-1]), flush=True)
        v.close()
        del sel
    del words, idx, out
    torch.cuda.empty_cache()
if args.out:
    with open(args.out, "w") as f:
        for ln in lines:
            f.write(json.dumps(ln) + "\n")
