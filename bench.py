#!/usr/bin/env python
"""bench.py — BASELINE.json config[1]: batched rank + select on a 1 GiB bit_vector, 1e8 uniform queries each.

One "step" = one pass of the hot path over one batch: 1e8 rank_1 queries followed by 1e8 select_1 queries against the
same 2^33-bit random bit vector (index resident in HBM).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--nbits-log2 33] [--queries 1e8]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N        (one rank per GPU)

N = 1  `value` = queries/s with the query and result arrays already in HBM (CUDA events on the launching stream);
       `e2e` = the same batch through the C ABI with pinned HOST buffers (H2D / kernel / D2H inside the timed region);
       `extras` = the rest of the metric: C5 count() on a 2^30 text, C4 wt.rank, C3 rrr / sd (tools/bench_extras.py).
N > 1  STRONG scaling of the same batch (north_star: "replicated index, NCCL all-gather of results only"): the 1e8 + 1e8
       queries are sharded over the N ranks through the C ABI's group calls (sdslgpu_group_rank / _select) and the
       results are all-gathered INSIDE the timed region, so every rank ends the step holding all 2e8 answers.
       `value` = 2e8 / step time (max over ranks) with the best gather this box supports, each one first checked to
       reproduce the single-GPU answers on every rank: packed peer stores (answers cross NVLink as 34-bit fields from the
       shards' last kernels), u64 peer stores, ncclAllGather; `variants` carries the others, the step without any
       gather, and the weak line (every rank answers a whole batch of its own).
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
sys.path.insert(0, os.path.join(ROOT, "tools"))
import __graft_entry__ as ge  # noqa: E402

# algorithmic bytes per query on the reference's layout
RANK_BYTES = 40    # 8 idx + 16 table pair + 8 data word read, 8 result written
SELECT_BYTES = 48  # 8 i + 8 superblock + 8 miniblock + 2*8 scanned words read (50 % density), 8 written
RESULT_BYTES = 8   # the written part of both: the "read-only" roofline variant leaves it out


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
        "higher_is_better": True, "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": workload_config(args, nbits, args.gpus),
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": cores, "kind": kind,
                         "sample": f"{sample} rank_1 + {sample} select_1 queries per step on the same 2^{args.nbits_log2}-bit vector",
                         "rank_qps": sample * args.steps / tr, "select_qps": sample * args.steps / ts},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def workload_config(args, nbits, world=1):
    cfg = {"workload": f"BASELINE config[1]: batched rank_1 + select_1 on a 2^{args.nbits_log2}-bit ({nbits // 8 / 2**30:g} GiB) random bit_vector, "
                       f"{args.queries:.0e} uniform rank + {args.queries:.0e} uniform select queries per step",
           "nbits": nbits, "rank_queries_per_step": int(args.queries), "select_queries_per_step": int(args.queries), "density": 0.5,
           "index": "replicated per GPU",
           "l2": "no flush needed: per step 1.6 GB of queries/results stream through and the 1.14 GiB index is gathered uniformly at random (L2 = 126 MB)"}
    if world > 1:
        cfg["queries"] = (f"ONE batch sharded over the {world} GPUs (strong scaling): rank r answers queries [r*n/{world}, (r+1)*n/{world}); "
                          "results all-gathered inside the timed region so every GPU holds all answers")
        cfg["parallelism"] = f"dp{world} (replicated index, sharded batch, all-gather of results)"
    else:
        cfg["queries"] = "one batch, one GPU"
    return cfg


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


def measured_traffic():
    """per-launch DRAM bytes (ncu dram__bytes_read.sum + dram__bytes_write.sum) of the pipelines, measured at this
    launch shape; the newest profiles/r*_traffic.json wins (its "session" key says which capture it is)"""
    import glob

    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_traffic.json")))
    if not files:
        return {}, None
    tj = json.load(open(files[-1]))
    return {k: v["dram_bytes_read"] + v["dram_bytes_write"] for k, v in tj.items() if isinstance(v, dict) and "dram_bytes_read" in v}, os.path.basename(files[-1])


def main():
    own_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nbits-log2", type=int, default=33)
    ap.add_argument("--queries", type=float, default=1e8)
    ap.add_argument("--ref-sample", type=float, default=1e8, help="queries per step and op for the CPU arms (default: the whole batch)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--extras", default="fm_count,wt_rank,rrr,sd", help="N = 1: which extra records to measure")
    ap.add_argument("--csa-log2", type=int, default=30)
    ap.add_argument("--c3-density", type=float, default=0.1)
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

    # the batch is the same on every rank (seed 0): N > 1 shards ONE batch
    words, idx, qr = make_workload(nbits, nq, 0)
    bv = pkg.BitVector(words, nbits, device=local)
    order = {"auto": pkg.ORDER_AUTO, "direct": pkg.ORDER_DIRECT, "binned": pkg.ORDER_BINNED}[args.order]
    bv.set_batch_order(order)
    index_bytes = (nbits // 224 + 1) * 32
    image_bytes = bv.device_bytes  # rank blocks + select samples, before the first large select batch adds its select sectors
    m = bv.arg_count(1)
    sel = qr.integers(1, m + 1, nq, dtype=np.uint64)

    h_idx = torch.from_numpy(idx.view(np.int64)).pin_memory()
    h_sel = torch.from_numpy(sel.view(np.int64)).pin_memory()
    d_idx, d_sel = h_idx.cuda(non_blocking=True), h_sel.cuda(non_blocking=True)
    group = sym = None
    if world > 1:
        ids = [pkg.group_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        group = pkg.Group.create_rank(ids[0], world, rank, local)
        sym = group.alloc(2 * nq * 8)  # symmetric result arrays: peers store straight into them
        d_out_r, d_out_s = sym.tensor(0)[:nq], sym.tensor(0)[nq:]
    else:
        d_out_r = torch.empty(nq, dtype=torch.int64, device="cuda")
        d_out_s = torch.empty(nq, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def plain_step(ev=None):
        if ev:
            ev[0].record()
        bv.rank(d_idx, 1, out=d_out_r)
        if ev:
            ev[1].record()
        bv.select(d_sel, 1, out=d_out_s)
        if ev:
            ev[2].record()

    def group_step(gather):
        def step(ev=None):
            cs = [torch.cuda.current_stream()]
            if ev:
                ev[0].record()
            group.rank([bv], 1, [d_idx], [d_out_r], gather=gather, streams=cs)
            if ev:
                ev[1].record()
            group.select([bv], 1, [d_sel], [d_out_s], gather=gather, streams=cs)
            if ev:
                ev[2].record()
        return step

    def timed(step, steps, warmup, clocks=False):
        """W untimed + exactly K timed steps, barrier + synchronize on both sides, device events, max over ranks"""
        for _ in range(warmup):
            step()
        barrier()
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
        sampler = ClockSampler(local) if clocks else None
        if sampler:
            sampler.start()
            time.sleep(0.25)
        t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        w0 = time.perf_counter()
        t_start.record()
        for k in range(steps):
            step(evs[k])
        t_end.record()
        barrier()
        w1 = time.perf_counter()
        ck = sampler.stop(w0, w1) if sampler else None
        total = t_start.elapsed_time(t_end) / steps
        a = sum(e[0].elapsed_time(e[1]) for e in evs) / steps
        b = sum(e[1].elapsed_time(e[2]) for e in evs) / steps
        mine = [total, a, b]
        per_rank = None
        if world > 1:
            t = torch.tensor(mine + [float((ck or {}).get("sm_mhz") or 0)], device="cuda", dtype=torch.float64)
            allr = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(allr, t)
            per_rank = [{"ms_per_step": x[0].item(), "rank_ms": x[1].item(), "select_ms": x[2].item(), "sm_mhz": int(x[3].item())} for x in allr]
            t3 = torch.tensor(mine, device="cuda", dtype=torch.float64)
            dist.all_reduce(t3, op=dist.ReduceOp.MAX)
            mine = t3.tolist()
        return {"ms_per_step": mine[0], "rank_ms": mine[1], "select_ms": mine[2], "clocks": ck, "per_rank": per_rank}

    variants = {}
    shard_q = nq // world
    per_call = shard_q if world > 1 else nq
    binned_r = args.order == "binned" or (args.order == "auto" and pkg.binned_wanted(index_bytes, per_call, select=False))
    binned_s = args.order == "binned" or (args.order == "auto" and pkg.binned_wanted(index_bytes, per_call, select=True))
    kernels_per_step = (3 if binned_r else 1) + (3 if binned_s else 1)
    if world == 1:
        main_t = timed(plain_step, args.steps, args.warmup, clocks=True)
        main_name = "single GPU"
        launches_per_step = kernels_per_step
    else:
        fused = group.fused_possible
        # this rank's own single-GPU answers for the whole batch: what every gathered array must equal
        ref_r, ref_s = torch.empty(nq, dtype=torch.int64, device="cuda"), torch.empty(nq, dtype=torch.int64, device="cuda")
        bv.rank(d_idx, 1, out=ref_r)
        bv.select(d_sel, 1, out=ref_s)

        def gathered_ok(gather):
            """one step with this gather on every rank; True iff every rank's arrays equal its single-GPU answers"""
            good = 1
            try:
                d_out_r.fill_(-3)
                d_out_s.fill_(-3)
                group_step(gather)()
                torch.cuda.synchronize()
                good = int(torch.equal(d_out_r, ref_r) and torch.equal(d_out_s, ref_s))
            except Exception as ex:  # an unsupported mode on this box
                print("bench.py: gather mode", gather, "failed:", repr(ex)[:200], file=sys.stderr)
                good = 0
            f = torch.tensor([good], device="cuda")
            dist.all_reduce(f, op=dist.ReduceOp.MIN)
            return bool(f.item())

        names = {pkg.GATHER_PACKED: "packed: the shards' last kernels store the answers as 34-bit fields into every peer's staging buffer over NVLink, each GPU widens what it received",
                 pkg.GATHER_FUSED: "fused: peer stores of u64 answers over NVLink from the last kernel of each shard",
                 pkg.GATHER_NCCL: "ncclAllGather after the shard kernels"}
        candidates = ([pkg.GATHER_PACKED, pkg.GATHER_FUSED] if fused else []) + [pkg.GATHER_NCCL]
        verified = [gm for gm in candidates if gathered_ok(gm)]
        if not verified:
            raise SystemExit("bench.py: no gather mode reproduces the single-GPU answers — refusing to report a number")
        # which of them is the fastest depends on N (packed pays an extra widening kernel, u64 pays NVLink bytes):
        # three untimed trial steps each, outside the timed region, decide
        trial = {gm: timed(group_step(gm), 3, 2)["ms_per_step"] for gm in verified}
        main_gather = min(trial, key=trial.get)
        main_name = names[main_gather]
        main_t = timed(group_step(main_gather), args.steps, args.warmup, clocks=True)
        # peer-store modes: 2 flag-exchange kernels per op (+ 1 widening kernel per op when packed)
        launches_per_step = kernels_per_step + (4 if main_gather != pkg.GATHER_NCCL else 0) + (2 if main_gather == pkg.GATHER_PACKED else 0)
        k2 = max(5, args.steps // 2)
        for gm, key in ((pkg.GATHER_PACKED, "packed_34bit"), (pkg.GATHER_FUSED, "fused_u64"), (pkg.GATHER_NCCL, "nccl_all_gather")):
            if gm != main_gather and gm in verified:
                variants[key] = timed(group_step(gm), k2, 3)
        variants["no_gather"] = timed(group_step(pkg.GATHER_NONE), k2, 3)
        variants["weak_whole_batch_per_rank"] = timed(plain_step, k2, 3)
        for k, v in variants.items():
            per = world * 2 * nq if k.startswith("weak") else 2 * nq
            v["value"] = per / (v["ms_per_step"] * 1e-3)
            v.pop("clocks", None)
        # leave the gathered results of the main variant in the arrays for the parity check below
        d_out_r.fill_(-3)
        d_out_s.fill_(-3)
        group_step(main_gather)()
        torch.cuda.synchronize()
        del ref_r, ref_s
    ms_per_step = main_t["ms_per_step"]
    rank_ms, sel_ms, clocks, per_rank = main_t["rank_ms"], main_t["select_ms"], main_t["clocks"], main_t["per_rank"]
    value = 2 * nq / (ms_per_step * 1e-3)

    # ---- e2e: the same step through the C ABI with HOST (pinned) buffers -------------------------------------------
    e2e = None
    if not args.no_e2e:
        lo, hi = (rank * shard_q, (rank + 1) * shard_q) if world > 1 else (0, nq)  # N > 1: every rank moves its shard
        n_loc = hi - lo
        h_out = torch.empty(max(n_loc, 1), dtype=torch.int64).pin_memory()
        hi_np, hs_np, ho_np = h_idx.numpy().view(np.uint64)[lo:hi], h_sel.numpy().view(np.uint64)[lo:hi], h_out.numpy().view(np.uint64)[:n_loc]

        def e2e_step():
            bv.rank(hi_np, 1, out=ho_np)
            bv.select(hs_np, 1, out=ho_np)

        def run_e2e(stepfn):
            stepfn()
            barrier()
            e0 = time.perf_counter()
            for _ in range(args.e2e_steps):
                stepfn()
            torch.cuda.synchronize()
            s_ = (time.perf_counter() - e0) / args.e2e_steps
            if world > 1:
                t = torch.tensor([s_], device="cuda", dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                s_ = t.item()
            return s_

        e2e_s = run_e2e(e2e_step)
        e2e = {"value": 2 * nq / e2e_s, "unit": "queries/s", "h2d_bytes_per_step": 2 * nq * 8, "d2h_bytes_per_step": 2 * nq * 8, "ms_per_step": e2e_s * 1e3,
               "path": "sdslgpu_rank / sdslgpu_select with pinned host buffers of u64 (chunked H2D / kernel / D2H)" +
                       ("" if world == 1 else f"; the batch is split over the {world} ranks, each moving its own shard")}
        if hasattr(pkg, "PackedBatch"):  # the int_vector<w> wire format (sdslgpu_rank_iv / _select_iv): w = bits needed, not 64
            try:
                pb = pkg.PackedBatch(bv, hi_np, hs_np, nbits)
                pk_s = run_e2e(pb.step)
                ok = pb.check(bv, hi_np[:100000], hs_np[:100000])
                e2e_u64 = dict(e2e)
                e2e = {"value": 2 * nq / pk_s, "unit": "queries/s", "h2d_bytes_per_step": pb.h2d_bytes * world, "d2h_bytes_per_step": pb.d2h_bytes * world,
                       "ms_per_step": pk_s * 1e3, "path": pb.describe() + ("" if world == 1 else f"; batch split over the {world} ranks"),
                       "values_identical_to_u64_path": ok, "u64_wire_format": e2e_u64}
            except Exception as ex:  # keep the u64 number
                e2e["packed_error"] = repr(ex)[:300]

    # ---- parity of what was timed (outside the timed region) + the CPU baseline ------------------------------------
    parity = cpu_baseline = None
    if world > 1:
        # every query of the gathered arrays against this rank's own single-GPU answers (device-side, all 2e8)
        chk = torch.empty(nq, dtype=torch.int64, device="cuda")
        bv.rank(d_idx, 1, out=chk)
        same_r = bool(torch.equal(chk, d_out_r))
        bv.select(d_sel, 1, out=chk)
        same_s = bool(torch.equal(chk, d_out_s))
        del chk
        flags = torch.tensor([int(same_r and same_s)], device="cuda")
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
        gathered_ok = bool(flags.item())
        if not gathered_ok:
            raise SystemExit("bench.py: gathered results differ from the single-GPU answers — refusing to report a number")
    if rank == 0:
        po = ge.load_oracle()
        if po.ref_available():
            chk, kind, cores = po.Ref().bv(words, nbits, with_select=True), "reference", os.cpu_count() or 1
        else:
            chk, kind, cores = po.Oracle().bv(words, nbits), "port", 1
        kw = {"threads": cores} if kind == "reference" else {}
        ns = nq if (kind == "reference" and not args.no_cpu_baseline) else min(nq, 200000)
        got_r = d_out_r[:ns].cpu().numpy().view(np.uint64)
        got_s = d_out_s[:ns].cpu().numpy().view(np.uint64)
        chk.rank(idx[: min(ns, 1000000)], 1, **kw)  # warm the threads / page in
        t0 = time.perf_counter()
        want_r = chk.rank(idx[:ns], 1, **kw)
        t1 = time.perf_counter()
        want_s = chk.select(sel[:ns], 1, **kw)
        t2 = time.perf_counter()
        ok = bool((got_r == want_r).all() and (got_s == want_s).all())
        parity = {"checked_queries": 2 * ns, "of": 2 * nq, "against": kind, "bit_exact": ok}
        if world > 1:
            parity["gathered_equals_single_gpu_on_all_queries_every_rank"] = gathered_ok
        if not ok:
            raise SystemExit("bench.py: GPU results differ from the reference — refusing to report a number")
        if not args.no_cpu_baseline and world == 1:
            cpu_baseline = {"value": 2 * ns / (t2 - t0), "unit": "queries/s", "cores": cores, "kind": kind,
                            "sample": f"{ns} rank_1 + {ns} select_1 queries: " + ("the whole timed batch" if ns == nq else "a prefix of the timed batch"),
                            "rank_qps": ns / (t1 - t0), "select_qps": ns / (t2 - t1)}
        del chk, got_r, got_s

    final_bytes = bv.device_bytes  # with the select sectors the timed select batches had built
    # ---- the rest of the metric (N = 1: C5 / C4 / C3 records; N > 1: C4 and C5 sharded + gathered) ------------------
    extras = {}
    if not args.no_extras:
        import bench_extras as bx

        po = ge.load_oracle()
        d_idx = d_sel = None
        if world == 1:
            d_out_r = d_out_s = None
            bv.close()
            torch.cuda.empty_cache()
            want = [x for x in args.extras.split(",") if x]
            for name, fn in (("fm_count", lambda: bx.fm_count_record(pkg, po, peak, peak_src, log2n=args.csa_log2)),
                             ("wt_rank", lambda: bx.wt_rank_record(pkg, po, peak, peak_src)),
                             ("c3", lambda: bx.compressed_records(pkg, po, peak, peak_src, log2n=args.nbits_log2, density=args.c3_density))):
                if name == "c3" and not ({"rrr", "sd"} & set(want)):
                    continue
                if name != "c3" and name not in want:
                    continue
                t0 = time.perf_counter()
                try:
                    r = fn()
                    if name == "c3":
                        for k in ("rrr", "sd"):
                            r[k]["wall_s"] = time.perf_counter() - t0
                            extras[k] = r[k]
                    else:
                        r["wall_s"] = time.perf_counter() - t0
                        extras[name] = r
                except Exception as ex:
                    extras[name] = {"error": repr(ex)[:400]}
                torch.cuda.empty_cache()
        else:
            extras = group_extras(pkg, po, group, sym, rank, world, local, args, peak, peak_src, barrier)

    if rank == 0:
        traffic, traffic_src = measured_traffic()
        if not (nbits == 1 << 33 and nq == int(1e8) and world == 1):
            traffic = {}
        k_rank, k_sel = ("bv_rank_kernel", "bv_select_kernel")
        n_rank, n_sel = "bv_rank_kernel<1,2>", "bv_select_kernel<1>"
        # binned: one op = three launches; the roofline entry is for the whole op (all three inside the event pair)
        if binned_r:
            k_rank, n_rank = "binned_rank_pipeline", "bin_tile_sort_kernel<1> + bin_apply_kernel<BvRankOp<1>> + bin_unsort_kernel"
        if binned_s:
            sect = final_bytes > image_bytes  # select sectors were built (include/sdslgpu.h, memory note of sdslgpu_select)
            k_sel = "binned_select_pipeline"
            n_sel = "bin_tile_sort_kernel<1> + bin_apply_kernel<%s<1>> + bin_unsort_kernel" % ("BvSelectSectOp" if sect else "BvSelectOp")
        per_gpu_q = shard_q if world > 1 else nq

        def roof(bytes_per_q, ms, kernel=None):
            a = per_gpu_q * bytes_per_q / (ms * 1e-3) / 1e9
            ro = per_gpu_q * (bytes_per_q - RESULT_BYTES) / (ms * 1e-3) / 1e9
            return {"bound": "hbm", "achieved": a, "peak": peak, "unit": "GB/s", "frac": a / peak, "frac_read_only": ro / peak, "traffic": traffic.get(kernel),
                    "traffic_source": traffic_src if traffic.get(kernel) else None, "peak_source": peak_src, "algorithmic_bytes_per_query": bytes_per_q,
                    "read_bytes_per_query": bytes_per_q - RESULT_BYTES, "kernel_ms": ms, "queries_per_launch": per_gpu_q}
        r_rank = dict(roof(RANK_BYTES, rank_ms, k_rank), kernel=n_rank, launches=3 if binned_r else 1, qps=per_gpu_q / (rank_ms * 1e-3))
        r_sel = dict(roof(SELECT_BYTES, sel_ms, k_sel), kernel=n_sel, launches=3 if binned_s else 1, qps=per_gpu_q / (sel_ms * 1e-3))
        dominant = r_sel if sel_ms >= rank_ms else r_rank
        cfg = workload_config(args, nbits, world)
        line = {
            "metric": "rank/select queries/s on 1 GiB bit_vector", "value": value, "unit": "queries/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": cfg,
            "roofline": dominant, "roofline_by_kernel": {"rank": r_rank, "select": r_sel},
            "cpu_baseline": cpu_baseline,
            "e2e": e2e,
            "gpu_launches": launches_per_step * args.steps, "batch_order": {"rank": "binned" if binned_r else "direct", "select": "binned" if binned_s else "direct"},
            "clocks": clocks, "per_rank": per_rank, "parity": parity,
            "index_device_bytes": final_bytes,
            "index_bytes_detail": {"bit_vector_raw": nbits // 8, "rank_blocks": index_bytes, "rank_blocks_and_select_samples": image_bytes,
                                   "select_sectors_built_by_the_first_large_select_batch": final_bytes - image_bytes,
                                   "note": "select answers from one 32-byte sector gather per query once the sectors exist; SDSLGPU_F_COMPACT handles keep the "
                                           "sampled select (1.58 instead of 1.26 ms per 1e8 queries here) and none of this memory"},
            "extras": extras,
        }
        if world > 1:
            bytes_per_answer = 34 / 8 if main_gather == pkg.GATHER_PACKED and nbits == 1 << 33 else 8
            nv = int((world - 1) * shard_q * bytes_per_answer * 2)
            line["gather"] = {"how": main_name, "nvlink_bytes_sent_per_rank_per_step": nv, "nvlink_bytes_received_per_rank_per_step": nv,
                              "received_gbs_per_rank": nv / (ms_per_step * 1e-3) / 1e9,
                              "limit": "each rank must RECEIVE (N-1)/N of the 2e8 answers per step over NVLink (900 GB/s per direction nominal; 8 B each as u64, 4.25 B packed): the answers, not the kernels, bound the gathered line"}
            line["value_no_gather"] = variants["no_gather"]["value"]
            line["variants"] = variants
        emit(line)
    if sym is not None:
        try:
            sym.release()
        except Exception:
            pass
    if group is not None:
        group.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def group_extras(pkg, po, group, sym, rank, world, local, args, peak, peak_src, barrier):
    """N > 1: BASELINE configs C4 (1e7 wt.rank, 1 -> 8 GPUs) and C5 (1e6 count(), 8 GPUs): index built on every rank,
    batch sharded, counts all-gathered inside the timed region; every rank's gathered array checked against its own
    single-GPU answers."""
    import torch
    import torch.distributed as dist

    import bench_extras as bx

    out = {}
    fused = group.fused_possible
    gather = pkg.GATHER_FUSED if fused else pkg.GATHER_NCCL

    def timed(fn, reps=7):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(reps):
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            t = torch.tensor([a.elapsed_time(b)], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ts.append(t.item())
        ts.sort()
        return ts[len(ts) // 2]

    cs = [torch.cuda.current_stream()]
    try:  # C4
        n, nq = 1 << 28, 10_000_000
        text = np.random.default_rng(42).integers(0, 256, n, dtype=np.uint8)
        wt = pkg.WtHuff(text, device=local)
        qr = np.random.default_rng(7)
        d_i = bx.dev(qr.integers(0, n + 1, nq, dtype=np.uint64))
        d_c = bx.dev(qr.integers(0, 256, nq, dtype=np.uint8))
        o = sym.tensor(0)[:nq]
        rec = {"config": f"C4 wt_huff<> on 2^28 uniform bytes, {nq} rank(i,c) sharded over {world} GPUs, results all-gathered", "unit": "queries/s"}
        for name, gm in (("gathered", gather), ("no_gather", pkg.GATHER_NONE)):
            ms = timed(lambda: group.wt_rank([wt], [d_i], [d_c], [o], gather=gm, streams=cs))
            rec[name] = {"ms": ms, "value": nq / (ms * 1e-3)}
        group.wt_rank([wt], [d_i], [d_c], [o], gather=gather, streams=cs)
        ok = torch.tensor([int(torch.equal(o, wt.rank(d_i, d_c)))], device="cuda")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        rec["value"], rec["gathered_equals_single_gpu_every_rank"] = rec["gathered"]["value"], bool(ok.item())
        out["wt_rank"] = rec
        wt.close()
        del text, d_i, d_c
        torch.cuda.empty_cache()
    except Exception as ex:
        out["wt_rank"] = {"error": repr(ex)[:400]}
    try:  # C5
        npat = 1_000_000
        text, flat, off = bx.c5_workload(args.csa_log2, npat)
        csa = pkg.CsaWt(text, device=local)
        del text
        d_f, d_o = bx.dev(flat), bx.dev(off)
        o = sym.tensor(0)[:npat]
        rec = {"config": f"C5 csa_wt<wt_huff<>> on a 2^{args.csa_log2}-byte text, {npat} count() |P|=20 sharded over {world} GPUs, counts all-gathered", "unit": "patterns/s",
               "index_device_bytes_per_gpu": csa.device_bytes}
        for name, gm in (("gathered", gather), ("no_gather", pkg.GATHER_NONE)):
            ms = timed(lambda: group.fm_count([csa], [d_f], [d_o], [o], gather=gm, streams=cs))
            rec[name] = {"ms": ms, "value": npat / (ms * 1e-3)}
        group.fm_count([csa], [d_f], [d_o], [o], gather=gather, streams=cs)
        ok = torch.tensor([int(torch.equal(o, csa.count(d_f, d_o)))], device="cuda")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        rec["value"], rec["gathered_equals_single_gpu_every_rank"] = rec["gathered"]["value"], bool(ok.item())
        out["fm_count"] = rec
        csa.close()
    except Exception as ex:
        out["fm_count"] = {"error": repr(ex)[:400]}
    return out


if __name__ == "__main__":
    main()
