#!/usr/bin/env python
"""rrr_vector<63> rank_1 / select_1 / select_0 at the C3 densities for the library named by $SDSLGPU_LIB (tools/variants.sh
builds -DRRR_SPARSE_PATH=0 as the A/B partner): 1e8 queries each on 2^33 bits, one JSON line; `sum_*` = wrapped sums of
the answers (equal across variants iff they answer alike)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import __graft_entry__ as ge  # noqa: E402
from bench_extras import gpu_random_words, time_gpu  # noqa: E402

pkg = ge.load_package()
nbits, nq = 1 << 33, 100_000_000
g = torch.Generator(device="cuda").manual_seed(3)
idx = torch.randint(0, nbits + 1, (nq,), dtype=torch.int64, device="cuda", generator=g)
out = torch.empty(nq, dtype=torch.int64, device="cuda")
res = {"variant": sys.argv[1] if len(sys.argv) > 1 else "product"}
for d in [float(x) for x in os.environ.get("PROBE_DENSITIES", "0.01,0.05,0.1,0.25,0.5,0.95").split(",")]:
    words = gpu_random_words(nbits, d, 42 + int(d * 100))
    v = pkg.RrrVector(words, nbits)
    del words
    tag = f"d{d:g}"
    res[tag + "_rank_ms"] = round(time_gpu(lambda: v.rank(idx, 1, out=out), 3), 4)
    res["sum_" + tag + "_rank"] = int(out.sum().item())
    for b in (1, 0):
        m = v.arg_count(b)
        sel = torch.randint(1, m + 1, (nq,), dtype=torch.int64, device="cuda", generator=g.manual_seed(11 + b))
        res[tag + f"_select{b}_ms"] = round(time_gpu(lambda: v.select(sel, b, out=out), 3), 4)
        res["sum_" + tag + f"_select{b}"] = int(out.sum().item())
        del sel
    v.close()
    torch.cuda.empty_cache()
print(json.dumps(res), flush=True)
