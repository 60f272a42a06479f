#!/usr/bin/env python
"""One line of timings for the library named by $SDSLGPU_LIB (tools/variants.sh run): plain rank / select of 1e8 queries
on 2^33 bits (the locality-ordered pipeline), select with and without select sectors (bv_device.cuh; run-time knob
SDSLGPU_SELECT_SECTORS), select_0, the direct-order select, other bin sizes for rank (VARIANT_CHUNKS="12,16"; the
library reads SDSLGPU_BIN_CHUNK_BYTES on every call), and sd_vector / rrr_vector ops at 50 % density.  `sum_*` = the
wrapped sum of all answers: equal across variants and knobs iff they answer alike."""
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


def timed(tag, fn):
    res[tag + "_ms"] = round(time_gpu(fn, 5), 4)
    res["sum_" + tag] = int(out.sum().item())


kinds = [("bv", pkg.BitVector), ("sd", pkg.SdVector), ("rrr", pkg.RrrVector)]
skip = os.environ.get("VARIANT_SKIP", "").split(",")
for name, cls in kinds:
    if name in skip:
        continue
    for sectors in ((0, 1) if name in ("bv", "sd") else (1,)):
        os.environ["SDSLGPU_SELECT_SECTORS"] = str(sectors)
        v = cls(words, nbits)
        tag = name if sectors else name + "_nosect"
        m = v.arg_count(1)
        sel = torch.randint(1, m + 1, (nq,), dtype=torch.int64, device="cuda", generator=g.manual_seed(7))
        b0 = v.device_bytes
        if sectors or name == "rrr":
            timed(tag + "_rank", lambda: v.rank(idx, 1, out=out))
        timed(tag + "_select", lambda: v.select(sel, 1, out=out))
        res[tag + "_bytes"] = [b0, v.device_bytes]
        if name == "bv":
            z = nbits - m
            sel0 = torch.randint(1, z + 1, (nq,), dtype=torch.int64, device="cuda", generator=g.manual_seed(8))
            timed(tag + "_select0", lambda: v.select(sel0, 0, out=out))
            del sel0
            v.set_batch_order(pkg.ORDER_DIRECT)
            timed(tag + "_select_direct", lambda: v.select(sel, 1, out=out))
            v.set_batch_order(pkg.ORDER_AUTO)
            if sectors:
                for mib in [int(x) for x in os.environ.get("VARIANT_CHUNKS", "").split(",") if x]:
                    os.environ["SDSLGPU_BIN_CHUNK_BYTES"] = str(mib << 20)
                    timed(f"bv_rank_chunk{mib}", lambda: v.rank(idx, 1, out=out))
                    timed(f"bv_select_chunk{mib}", lambda: v.select(sel, 1, out=out))
                os.environ.pop("SDSLGPU_BIN_CHUNK_BYTES", None)
        v.close()
        del sel
os.environ.pop("SDSLGPU_SELECT_SECTORS", None)
if "bv" not in skip:
    for stride in [int(x) for x in os.environ.get("VARIANT_STRIDES", "").split(",") if x]:  # B-bits per select sector
        os.environ["SDSLGPU_SELECT_SECTOR_STRIDE"] = str(stride)
        v = pkg.BitVector(words, nbits)
        sel = torch.randint(1, v.arg_count(1) + 1, (nq,), dtype=torch.int64, device="cuda", generator=g.manual_seed(7))
        timed(f"bv_select_stride{stride}", lambda: v.select(sel, 1, out=out))
        res[f"bv_stride{stride}_bytes"] = v.device_bytes
        v.close()
        del sel
    os.environ.pop("SDSLGPU_SELECT_SECTOR_STRIDE", None)
print(json.dumps(res), flush=True)
