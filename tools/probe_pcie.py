#!/usr/bin/env python
"""Which host resource caps the end-to-end (host-buffer) numbers at N > 1?  Every rank copies pinned host memory to its
GPU and back, alone and all ranks at once, one direction and both; prints per-rank and aggregate GB/s.

    python tools/probe_pcie.py                                            (N = 1)
    python -m torch.distributed.run --nproc-per-node N ... tools/probe_pcie.py
"""
import json
import os
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
nbytes = 1 << 30
h_in = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
h_out = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
h_in.fill_(1)
d_a = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
d_b = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def run(mode, reps=4):
    def once():
        if mode in ("h2d", "both"):
            with torch.cuda.stream(s1):
                d_a.copy_(h_in, non_blocking=True)
        if mode in ("d2h", "both"):
            with torch.cuda.stream(s2):
                h_out.copy_(d_b, non_blocking=True)
    once()
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        once()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    gbs = reps * nbytes * (2 if mode == "both" else 1) / dt / 1e9
    if world > 1:
        t = torch.tensor([gbs], device="cuda", dtype=torch.float64)
        allr = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allr, t)
        return [x.item() for x in allr]
    return [gbs]


numa = None
try:
    bus = torch.cuda.get_device_properties(local).pci_bus_id
    numa = open(f"/sys/bus/pci/devices/0000:{bus:02x}:00.0/numa_node").read().strip()
except Exception:
    pass
res = {"n_gpus": world}
for mode in ("h2d", "d2h", "both"):
    per = run(mode)
    res[mode] = {"per_rank_gbs": [round(x, 1) for x in per], "aggregate_gbs": round(sum(per), 1)}
if rank == 0:
    res["host_cores"] = os.cpu_count()
    res["numa_node_of_gpu0"] = numa
    print(json.dumps(res), flush=True)
if world > 1:
    dist.destroy_process_group()
