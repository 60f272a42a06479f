#!/usr/bin/env python
"""bench.py — BASELINE.json config[1]: batched rank + select on a 1 GiB bit_vector, 1e8 uniform queries each.

One "step" = one pass of the hot path over one batch: 1e8 rank_1 queries followed by 1e8 select_1 queries
against the same 2^33-bit random bit vector (index resident in HBM).  `value` = queries/s with the query and
result arrays already in HBM (CUDA events on the launching stream, max over ranks); `e2e` = the same batch
through the C-ABI with pinned HOST buffers (chunked H2D / kernel / D2H pipeline inside the timed region).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--nbits-log2 33] [--queries 1e8]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N        (one rank per GPU)

N > 1: the index is replicated, every rank answers its own shard of queries (weak scaling: 1e8 + 1e8 per
rank), no data-path collective; timing is barrier + device events, max over ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

# algorithmic bytes per query on the reference's layout
RANK_BYTES = 40    # 8 idx + 16 table pair + 8 data word read, 8 result written
SELECT_BYTES = 48  # 8 i + 8 superblock + 8 miniblock + 2*8 scanned words read (50 % density), 8 written


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)"""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.gpu)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for (t, r) in self.rows if t0 <= t <= t1 + 0.12] or [r for (_, r) in self.rows[-3:]]
        sm = sorted(int(float(r[1])) for r in rows if len(r) > 2)
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            for k, nm in enumerate(names):
                if len(r) > 5 + k and r[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        pw = [float(r[3]) for r in rows if len(r) > 3 and r[3].replace(".", "", 1).isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_min_mhz": sm[0] if sm else None, "sm_max_mhz": int(float(rows[0][2])) if rows else None,
                "reasons": sorted(reasons), "samples": len(rows), "power_w_max": max(pw) if pw else None}


def make_workload(nbits, nq, rank_seed):
    """words = successive rng draws (util::set_random_bits semantics), uniform queries"""
    rng = np.random.default_rng(42)
    words = rng.integers(0, 2**64, (nbits + 63) // 64, dtype=np.uint64)
    qr = np.random.default_rng(7 + 1000 * rank_seed)
    idx = qr.integers(0, nbits + 1, nq, dtype=np.uint64)
    return words, idx, qr


def run_reference(args, rank, world):
    """--impl reference: the UNMODIFIED reference (oracle/_ref) on the host cores; bounded sample per step."""
    if rank != 0:
        return
    po = ge.load_oracle()
    nbits = 1 << args.nbits_log2
    cores = os.cpu_count() or 1
    sample = int(min(args.queries, args.ref_sample))
    words, idx, qr = make_workload(nbits, sample, 0)
    kind = "reference" if po.ref_available() else "port"
    if kind == "reference":
        h = po.Ref().bv(words, nbits, with_select=True)
        run_r = lambda: h.rank(idx, 1, threads=cores)
        m = int(h.rank([nbits], 1)[0])
        sel = qr.integers(1, m + 1, sample, dtype=np.uint64)
        run_s = lambda: h.select(sel, 1, threads=cores)
    else:
        h = po.Oracle().bv(words, nbits)
        cores = 1
        run_r = lambda: h.rank(idx, 1)
        m = int(h.rank([nbits], 1)[0])
        sel = qr.integers(1, m + 1, sample, dtype=np.uint64)
        run_s = lambda: h.select(sel, 1)
    for _ in range(args.warmup):
        run_r(), run_s()
    tr = ts = 0.0
    for _ in range(args.steps):
        t0 = time.perf_counter(); run_r(); t1 = time.perf_counter(); run_s(); t2 = time.perf_counter()
        tr += t1 - t0
        ts += t2 - t1
    qps = 2 * sample * args.steps / (tr + ts)
    line = {
        "impl": "reference", "metric": "rank/select queries/s on 1 GiB bit_vector", "value": qps, "unit": "queries/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * (tr + ts) / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": workload_config(args, nbits),
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": cores, "kind": kind,
                         "sample": f"{sample} rank_1 + {sample} select_1 queries per step on the same 2^{args.nbits_log2}-bit vector",
                         "rank_qps": sample * args.steps / tr, "select_qps": sample * args.steps / ts},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def workload_config(args, nbits):
    return {"workload": f"BASELINE config[1]: batched rank_1 + select_1 on a 2^{args.nbits_log2}-bit ({nbits // 8 / 2**30:g} GiB) random bit_vector, "
                        f"{args.queries:.0e} uniform rank + {args.queries:.0e} uniform select queries per step per GPU",
            "nbits": nbits, "rank_queries_per_step": int(args.queries), "select_queries_per_step": int(args.queries),
            "density": 0.5, "index": "replicated per GPU", "queries": "sharded (independent per rank), no data-path collective",
            "l2": "no flush needed: per step 1.6 GB of queries/results stream through and the 1.14 GiB index is gathered uniformly at random (L2 = 126 MB)"}


_REAL_STDOUT = None


def own_stdout():
    """stdout carries exactly ONE JSON line: everything else any library prints there (NCCL's version banner, ...)
    is sent to stderr by pointing fd 1 at fd 2 for the whole run; emit() writes the line to the real stdout."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    own_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nbits-log2", type=int, default=33)
    ap.add_argument("--queries", type=float, default=1e8)
    ap.add_argument("--ref-sample", type=float, default=2e7, help="queries per step for the CPU arms")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--order", default="auto", choices=["auto", "direct", "binned"],
                    help="sdslgpu_set_batch_order: auto picks the binned pipeline for this workload")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = ge.load_package()
    nbits = 1 << args.nbits_log2
    nq = int(args.queries)
    peak, peak_src = peaks()

    words, idx, qr = make_workload(nbits, nq, rank)
    bv = pkg.BitVector(words, nbits, device=local)
    bv.set_batch_order({"auto": pkg.ORDER_AUTO, "direct": pkg.ORDER_DIRECT, "binned": pkg.ORDER_BINNED}[args.order])
    # what AUTO resolves to (include/sdslgpu.h): binned iff index >= 192 MB, n >= 2^21 and n >= index_bytes / 64
    index_bytes = (nbits // 224 + 1) * 32
    binned = args.order == "binned" or (args.order == "auto" and index_bytes >= 192 << 20 and nq >= 1 << 21 and nq >= index_bytes // 64)
    m = bv.arg_count(1)
    sel = qr.integers(1, m + 1, nq, dtype=np.uint64)

    # pinned host buffers for the e2e leg; device-resident copies for the kernel-only leg
    h_idx = torch.from_numpy(idx.view(np.int64)).pin_memory()
    h_sel = torch.from_numpy(sel.view(np.int64)).pin_memory()
    h_out = torch.empty(nq, dtype=torch.int64).pin_memory()
    d_idx, d_sel = h_idx.cuda(non_blocking=True), h_sel.cuda(non_blocking=True)
    d_out_r = torch.empty(nq, dtype=torch.int64, device="cuda")
    d_out_s = torch.empty(nq, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()

    def step(ev=None):
        if ev:
            ev[0].record()
        bv.rank(d_idx, 1, out=d_out_r)
        if ev:
            ev[1].record()
        bv.select(d_sel, 1, out=d_out_s)
        if ev:
            ev[2].record()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    sampler = ClockSampler(local)  # every rank watches its own GPU
    sampler.start()
    time.sleep(0.25)
    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    barrier()
    w0 = time.perf_counter()
    t_start.record()
    for k in range(args.steps):
        step(evs[k])
    t_end.record()
    barrier()
    w1 = time.perf_counter()
    clocks = sampler.stop(w0, w1)
    total_ms = t_start.elapsed_time(t_end)
    rank_ms = sum(e[0].elapsed_time(e[1]) for e in evs) / args.steps
    sel_ms = sum(e[1].elapsed_time(e[2]) for e in evs) / args.steps
    per_rank = None
    if world > 1:
        mine = torch.tensor([total_ms / args.steps, rank_ms, sel_ms, float(clocks.get("sm_mhz") or 0)], device="cuda", dtype=torch.float64)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = [{"ms_per_step": a[0].item(), "rank_ms": a[1].item(), "select_ms": a[2].item(), "sm_mhz": int(a[3].item())} for a in allr]
        t = torch.tensor([total_ms, rank_ms, sel_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, rank_ms, sel_ms = t.tolist()
    ms_per_step = total_ms / args.steps
    value = world * 2 * nq / (ms_per_step * 1e-3)

    # e2e: the same step through the C-ABI with HOST (pinned) buffers
    def e2e_step():
        bv.rank(h_idx.numpy().view(np.uint64), 1, out=h_out.numpy().view(np.uint64))
        bv.select(h_sel.numpy().view(np.uint64), 1, out=h_out.numpy().view(np.uint64))

    e2e_step()
    barrier()
    e0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - e0) / args.e2e_steps
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = t.item()
    e2e_val = world * 2 * nq / e2e_s

    # parity spot check of what was timed (outside the timed region): the reference / oracle on a sample
    parity = None
    cpu_baseline = None
    if rank == 0:
        po = ge.load_oracle()
        ns = 200000
        got_r = d_out_r[:ns].cpu().numpy().view(np.uint64)
        got_s = d_out_s[:ns].cpu().numpy().view(np.uint64)
        if po.ref_available():
            chk = po.Ref().bv(words, nbits, with_select=True)
            kind, cores = "reference", os.cpu_count() or 1
        else:
            chk = po.Oracle().bv(words, nbits)
            kind, cores = "port", 1
        ok = bool((got_r == chk.rank(idx[:ns], 1)).all() and (got_s == chk.select(sel[:ns], 1)).all())
        parity = {"checked_queries": 2 * ns, "against": kind, "bit_exact": ok}
        if not ok:
            raise SystemExit("bench.py: GPU results differ from the reference — refusing to report a number")
        if not args.no_cpu_baseline and world == 1:  # the CPU baseline is reported at N = 1 only
            sample = int(min(nq, args.ref_sample))
            kw = {"threads": cores} if kind == "reference" else {}
            chk.rank(idx[: sample // 10], 1, **kw)
            t0 = time.perf_counter(); chk.rank(idx[:sample], 1, **kw); t1 = time.perf_counter()
            chk.select(sel[:sample], 1, **kw); t2 = time.perf_counter()
            cpu_baseline = {"value": 2 * sample / (t2 - t0), "unit": "queries/s", "cores": cores, "kind": kind,
                            "sample": f"{sample} rank_1 + {sample} select_1 queries of the same batch on the same vector",
                            "rank_qps": sample / (t1 - t0), "select_qps": sample / (t2 - t1)}

    if rank == 0:
        traffic = {}
        tp = os.path.join(ROOT, "profiles", "r01_traffic.json")
        if os.path.exists(tp) and nbits == 1 << 33 and nq == int(1e8):  # measured for exactly this launch shape
            tj = json.load(open(tp))
            traffic = {k: v["dram_bytes_read"] + v["dram_bytes_write"] for k, v in tj.items() if isinstance(v, dict)}
        k_rank, k_sel = ("bv_rank_kernel", "bv_select_kernel")
        n_rank, n_sel = "bv_rank_kernel<1,2>", "bv_select_kernel<1>"
        if binned:  # one op = three launches; the roofline entry is for the whole op (all three inside the event pair)
            k_rank, k_sel = "binned_rank_pipeline", "binned_select_pipeline"
            n_rank = "bin_tile_sort_kernel<1> + bin_apply_kernel<BvRankOp<1>> + bin_unsort_kernel"
            n_sel = "bin_tile_sort_kernel<1> + bin_apply_kernel<BvSelectOp<1>> + bin_unsort_kernel"

        def roof(bytes_per_q, ms, kernel=None):
            a = nq * bytes_per_q / (ms * 1e-3) / 1e9
            return {"bound": "hbm", "achieved": a, "peak": peak, "unit": "GB/s", "frac": a / peak, "traffic": traffic.get(kernel),
                    "peak_source": peak_src, "algorithmic_bytes_per_query": bytes_per_q, "kernel_ms": ms}
        r_rank = dict(roof(RANK_BYTES, rank_ms, k_rank), kernel=n_rank, launches=3 if binned else 1, qps=nq / (rank_ms * 1e-3))
        r_sel = dict(roof(SELECT_BYTES, sel_ms, k_sel), kernel=n_sel, launches=3 if binned else 1, qps=nq / (sel_ms * 1e-3))
        dominant = r_sel if sel_ms >= rank_ms else r_rank
        line = {
            "metric": "rank/select queries/s on 1 GiB bit_vector", "value": value, "unit": "queries/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": workload_config(args, nbits),
            "roofline": dominant, "roofline_by_kernel": {"rank": r_rank, "select": r_sel},
            "cpu_baseline": cpu_baseline,
            "e2e": {"value": e2e_val, "unit": "queries/s", "h2d_bytes_per_step": 2 * nq * 8, "d2h_bytes_per_step": 2 * nq * 8,
                    "ms_per_step": e2e_s * 1e3, "path": "sdslgpu_rank/sdslgpu_select with pinned host buffers (chunked H2D/kernel/D2H)"},
            "gpu_launches": (6 if binned else 2) * args.steps, "batch_order": "binned" if binned else "direct",
            "clocks": clocks, "per_rank": per_rank, "parity": parity,
            "index_device_bytes": bv.device_bytes,
        }
        emit(line)
    bv.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
