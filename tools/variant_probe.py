#!/usr/bin/env python
"""One line of timings for the library named by $SDSLGPU_LIB: plain rank / select (binned) and rrr rank / select at
50 % density, 1e8 queries each on 2^33 bits (tools/variants.sh run)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import __graft_entry__ as ge  # noqa: E402
from bench_extras import time_gpu  # noqa: E402

pkg = ge.load_package()
nbits, nq = 1 << 33, 100_000_000
g = torch.Generator(device="cuda").manual_seed(42)
words = torch.randint(-(2**63), 2**63 - 1, (nbits // 64,), dtype=torch.int64, device="cuda", generator=g)
idx = torch.randint(0, nbits + 1, (nq,), dtype=torch.int64, device="cuda", generator=g)
out = torch.empty(nq, dtype=torch.int64, device="cuda")
res = {"variant": sys.argv[1] if len(sys.argv) > 1 else os.environ.get("SDSLGPU_LIB", "product")}
for name, cls in (("bv", pkg.BitVector), ("rrr", pkg.RrrVector)):
    v = cls(words, nbits)
    m = v.arg_count(1)
    sel = torch.randint(1, m + 1, (nq,), dtype=torch.int64, device="cuda", generator=g)
    res[name + "_rank_ms"] = round(time_gpu(lambda: v.rank(idx, 1, out=out), 5), 4)
    res[name + "_select_ms"] = round(time_gpu(lambda: v.select(sel, 1, out=out), 5), 4)
    v.close()
    del sel
print(json.dumps(res), flush=True)
