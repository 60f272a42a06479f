#!/usr/bin/env python
"""tools/bench_all.py — every BASELINE.json config on one B200, one JSON line per (config, operation):

    python tools/bench_all.py [--configs C2,C3,C4,C5] [--csa-log2 28] [--reps 5] [--out gpurun_out/bench_all.jsonl]

For each operation: device-resident queries, CUDA-event timing of the kernel(s) (median of --reps after 2 warm-ups),
queries/s, achieved ALGORITHMIC GB/s (bytes per query from SURVEY.md §8(d)) and its fraction of the measured HBM peak,
a bit-exactness check of a sample against the unmodified reference (oracle/_ref), and the reference's own CPU rate
on a bounded sample with all host threads.  bench.py stays the single-line headline (config[1]); this script is
the evidence for the other rows.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402
import torch  # noqa: E402

pkg = ge.load_package()
po = ge.load_oracle()
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
CORES = os.cpu_count() or 1
REF = po.Ref() if po.ref_available() else None
OUT = None


def emit(rec):
    rec["peak_gbs"] = PEAK
    line = json.dumps(rec)
    print(line, flush=True)
    if OUT:
        OUT.write(line + "\n")
        OUT.flush()


def dev(a):
    if a.dtype == np.uint64:
        return torch.from_numpy(a.view(np.int64)).cuda()
    return torch.from_numpy(a).cuda()


def host(t):
    a = t.cpu().numpy()
    return a.view(np.uint64) if a.dtype == np.int64 else a


def time_gpu(fn, reps):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def time_cpu(fn):
    t0 = time.perf_counter()
    r = fn()
    return time.perf_counter() - t0, r


def record(cfg, op, nq, ms, best_ms, bytes_per_q, parity, cpu=None, extra=None):
    qps = nq / (ms * 1e-3)
    gbs = qps * bytes_per_q / 1e9
    rec = {"config": cfg, "op": op, "queries": nq, "ms": ms, "best_ms": best_ms, "qps": qps, "algorithmic_bytes_per_query": bytes_per_q,
           "achieved_gbs": gbs, "frac_of_hbm_peak": gbs / PEAK, "parity": parity}
    if cpu:
        rec["cpu_reference"] = cpu
        rec["speedup_vs_cpu"] = qps / cpu["qps"]
    if extra:
        rec.update(extra)
    emit(rec)


def gpu_random_words(nbits, density, seed):
    """Bernoulli(density) bit vector generated on the device -> int64 words tensor"""
    g = torch.Generator(device="cuda").manual_seed(seed)
    nw = (nbits + 63) // 64
    out = torch.empty(nw, dtype=torch.int64, device="cuda")
    chunk = 1 << 21  # words per chunk
    w = (torch.ones(64, dtype=torch.int64, device="cuda") << torch.arange(64, device="cuda", dtype=torch.int64))
    for lo in range(0, nw, chunk):
        hi = min(nw, lo + chunk)
        bits = (torch.rand((hi - lo, 64), device="cuda", generator=g) < density).to(torch.int64)
        out[lo:hi] = (bits * w).sum(1)
    if nbits % 64:
        out[-1] &= (1 << (nbits % 64)) - 1
    return out


# --------------------------------------------------------------------------------------------------------
def run_c2(args):
    nbits, nq = 1 << args.nbits_log2, int(args.queries)
    rng = np.random.default_rng(42)
    words = rng.integers(0, 2**64, nbits // 64, dtype=np.uint64)
    qr = np.random.default_rng(7)
    idx = qr.integers(0, nbits + 1, nq, dtype=np.uint64)
    ref = REF.bv(words, nbits, with_select=True) if REF else None
    ns = min(nq, int(args.cpu_sample))
    for layout, flags in (("b200", 0), ("sdsl", pkg.F_SDSL_LAYOUT)):
        bv = pkg.BitVector(words, nbits, flags=flags)
        d_idx, d_out = dev(idx), torch.empty(nq, dtype=torch.int64, device="cuda")
        orders = (("binned", pkg.ORDER_BINNED), ("direct", pkg.ORDER_DIRECT)) if layout == "b200" else (("direct", pkg.ORDER_DIRECT),)
        for b in (1, 0):
            par, cpu = None, None
            for oname, order in orders:
                bv.set_batch_order(order)
                ms, best = time_gpu(lambda: bv.rank(d_idx, b, out=d_out), args.reps)
                if ref:
                    if cpu is None:
                        t, want = time_cpu(lambda: ref.rank(idx[:ns], b, threads=CORES))
                        cpu = {"qps": ns / t, "cores": CORES, "sample": ns}
                    par = bool((host(d_out[:ns]) == want).all())
                record(f"C2 2^{args.nbits_log2}-bit random bit_vector", f"rank_{b} [{layout} layout, {oname} order]", nq, ms, best, 40, par, cpu,
                       {"index_bytes": bv.device_bytes, "batch_order": oname})
        if layout == "b200":
            for b in (1, 0):
                m = bv.arg_count(b)
                sel = qr.integers(1, m + 1, nq, dtype=np.uint64)
                d_sel = dev(sel)
                par, cpu = None, None
                for oname, order in orders:
                    bv.set_batch_order(order)
                    ms, best = time_gpu(lambda: bv.select(d_sel, b, out=d_out), args.reps)
                    if ref:
                        if cpu is None:
                            t, want = time_cpu(lambda: ref.select(sel[:ns], b, threads=CORES))
                            cpu = {"qps": ns / t, "cores": CORES, "sample": ns}
                        par = bool((host(d_out[:ns]) == want).all())
                    record(f"C2 2^{args.nbits_log2}-bit random bit_vector", f"select_{b} [{oname} order]", nq, ms, best, 48, par, cpu, {"batch_order": oname})
            bv.set_batch_order(pkg.ORDER_DIRECT)
            ms, best = time_gpu(lambda: bv.access(d_idx, out=d_out), args.reps)
            record(f"C2 2^{args.nbits_log2}-bit random bit_vector", "access", nq, ms, best, 24, None)
            # the same batch in ascending order: neighbouring queries share sectors, the kernel is unchanged
            d_sorted = torch.sort(d_idx.view(torch.int64)).values
            d_chk = torch.empty_like(d_out)
            bv.rank(d_idx, 1, out=d_chk)
            want_sorted = torch.sort(d_chk).values  # rank is monotone in i
            ms, best = time_gpu(lambda: bv.rank(d_sorted, 1, out=d_out), args.reps)
            record(f"C2 2^{args.nbits_log2}-bit random bit_vector", "rank_1, queries sorted ascending", nq, ms, best, 40,
                   bool(torch.equal(d_out, want_sorted)), None)
            del d_sorted, d_chk, want_sorted
        bv.close()
        del d_idx, d_out
    torch.cuda.empty_cache()


def run_c3(args):
    nbits = 1 << args.nbits_log2
    nq = int(args.queries_c3)
    ns = min(nq, int(args.cpu_sample) // 4)
    for d in [float(x) for x in args.densities.split(",")]:
        words_d = gpu_random_words(nbits, d, 42 + int(d * 100))
        plain = pkg.BitVector(words_d, nbits)
        qr = np.random.default_rng(7)
        idx = qr.integers(0, nbits + 1, nq, dtype=np.uint64)
        m = plain.arg_count(1)
        sel = qr.integers(1, m + 1, nq, dtype=np.uint64)
        d_idx, d_sel = dev(idx), dev(sel)
        want_rank, want_sel = plain.rank(d_idx, 1), plain.select(d_sel, 1)
        words_h = host(words_d) if (REF and d in args.cpu_densities) else None
        for kind, cls, rank_bytes, sel_bytes in (("rrr_vector<63>", pkg.RrrVector, 76, 76), ("sd_vector<>", pkg.SdVector, 72, 56)):
            t0 = time.perf_counter()
            v = cls(words_d, nbits)
            torch.cuda.synchronize()
            build_s = time.perf_counter() - t0
            d_out = torch.empty(nq, dtype=torch.int64, device="cuda")
            ref = None
            if words_h is not None:
                tb, ref = time_cpu(lambda: (REF.rrr if kind.startswith("rrr") else REF.sd)(words_h, nbits))
            for op, q, dq, want, nbytes in (("rank_1", idx, d_idx, want_rank, rank_bytes), ("select_1", sel, d_sel, want_sel, sel_bytes)):
                fn = (lambda: v.rank(dq, 1, out=d_out)) if op == "rank_1" else (lambda: v.select(dq, 1, out=d_out))
                cpu, w = None, None
                for oname, order in (("auto", pkg.ORDER_AUTO), ("direct", pkg.ORDER_DIRECT)):
                    if oname == "direct" and kind.startswith("rrr") and op == "select_1":
                        continue  # rrr select has no binned form: auto == direct
                    v.set_batch_order(order)
                    ms, best = time_gpu(fn, args.reps)
                    par = bool((d_out == want).all().item())  # vs the plain-vector kernels (themselves reference-checked in C2)
                    if ref is not None:
                        if cpu is None:
                            f = (lambda: ref.rank(q[:ns], 1, threads=CORES)) if op == "rank_1" else (lambda: ref.select(q[:ns], 1, threads=CORES))
                            t, w = time_cpu(f)
                            cpu = {"qps": ns / t, "cores": CORES, "sample": ns, "build_s": tb}
                        par = par and bool((host(d_out[:ns]) == w).all())
                    record(f"C3 2^{args.nbits_log2} bits, density {d:g}", f"{kind} {op} [{oname} order]", nq, ms, best, nbytes, par, cpu,
                           {"index_bytes": v.device_bytes, "gpu_build_s": build_s, "bits_per_bit": 8.0 * v.device_bytes / nbits, "batch_order": oname})
            v.close()
            del ref
        plain.close()
        del words_d, d_idx, d_sel, want_rank, want_sel
        torch.cuda.empty_cache()


def run_c4(args):
    n, nq = 1 << args.text_log2, int(args.queries_c4)
    rng = np.random.default_rng(42)
    text = rng.integers(0, 256, n, dtype=np.uint8)
    t0 = time.perf_counter()
    wt = pkg.WtHuff(text)
    build_s = time.perf_counter() - t0
    qr = np.random.default_rng(7)
    i = qr.integers(0, n + 1, nq, dtype=np.uint64)
    c = qr.integers(0, 256, nq, dtype=np.uint8)
    d_i, d_c, d_out = dev(i), dev(c), torch.empty(nq, dtype=torch.int64, device="cuda")
    ref, ref_build = None, None
    if REF:
        ref_build, ref = time_cpu(lambda: REF.wt_huff(text))
    ns = min(nq, int(args.cpu_sample) // 8)
    cfg = f"C4 wt_huff on 2^{args.text_log2} uniform bytes (sigma 256)"
    ms, best = time_gpu(lambda: wt.rank(d_i, d_c, out=d_out), args.reps)
    par, cpu = None, None
    if ref:
        t, w = time_cpu(lambda: ref.rank(i[:ns], c[:ns], threads=CORES))
        par, cpu = bool((host(d_out[:ns]) == w).all()), {"qps": ns / t, "cores": CORES, "sample": ns, "build_s": ref_build}
    record(cfg, "wt.rank(i,c)", nq, ms, best, 209, par, cpu, {"index_bytes": wt.device_bytes, "build_s": build_s})
    j = qr.integers(0, n, nq, dtype=np.uint64)
    d_j = dev(j)
    ms, best = time_gpu(lambda: wt.inverse_select(d_j), args.reps)
    rnk, sym = wt.inverse_select(d_j)
    par = bool((host(sym).astype(np.uint8) == text[j.astype(np.int64)]).all())
    record(cfg, "wt.inverse_select(i)", nq, ms, best, 8 + 8 * 24 + 16, par)
    occ = host(wt.rank(torch.full((256,), n, dtype=torch.int64, device="cuda"), torch.arange(256, dtype=torch.uint8, device="cuda")))
    k = (qr.integers(0, 2**62, nq, dtype=np.uint64) % occ[c.astype(np.int64)]) + np.uint64(1)
    d_k = dev(k)
    ms, best = time_gpu(lambda: wt.select(d_k, d_c, out=d_out), args.reps)
    par, cpu = None, None
    if ref:
        t, w = time_cpu(lambda: ref.select(k[:ns], c[:ns], threads=CORES))
        par, cpu = bool((host(d_out[:ns]) == w).all()), {"qps": ns / t, "cores": CORES, "sample": ns}
    record(cfg, "wt.select(i,c)", nq, ms, best, 9 + 8 * 48 + 8, par, cpu)
    wt.close()
    torch.cuda.empty_cache()


def run_c5(args):
    n, npat, plen = 1 << args.csa_log2, int(args.patterns), 20
    rng = np.random.default_rng(42)
    text = rng.integers(1, 256, n, dtype=np.uint8)
    t0 = time.perf_counter()
    csa = pkg.CsaWt(text)
    build_s = time.perf_counter() - t0
    qr = np.random.default_rng(7)
    starts = qr.integers(0, n - plen, npat)
    flat = text[(starts[:, None] + np.arange(plen)[None, :])].reshape(-1).copy()
    off = (np.arange(npat + 1, dtype=np.uint64) * np.uint64(plen))
    d_flat, d_off = dev(flat), dev(off)
    cfg = f"C5 csa_wt<wt_huff<>> on 2^{args.csa_log2} uniform bytes 1..255, {npat} patterns |P|=20 sampled from the text"
    ref, ref_build = None, None
    if REF and args.csa_ref:
        ref_build, ref = time_cpu(lambda: REF.csa(text))
    ns = min(npat, 200000)
    ms, best = time_gpu(lambda: csa.count(d_flat, d_off), args.reps)
    cnt = host(csa.count(d_flat, d_off))
    par, cpu = bool((cnt >= 1).all()), None
    if ref:
        t, w = time_cpu(lambda: ref.count(flat[: ns * plen], off[: ns + 1], threads=CORES))
        par, cpu = par and bool((cnt[:ns] == w).all()), {"qps": ns / t, "cores": CORES, "sample": ns, "build_s": ref_build}
    record(cfg, "count()", npat, ms, best, 19 * 2 * 32, par, cpu, {"index_bytes": csa.device_bytes, "build_s": build_s, "unit": "patterns/s"})
    ms, best = time_gpu(lambda: csa.locate(d_flat, d_off), args.reps)
    occ_off, occ = csa.locate(d_flat, d_off)
    tot = int(occ.numel())
    par, cpu = None, None
    if ref:
        nl = min(ns, 50000)
        t, w = time_cpu(lambda: ref.locate(flat[: nl * plen], off[: nl + 1], threads=CORES))
        oo, oc = host(occ_off), host(occ)
        par = bool((oo[: nl + 1] == w[0]).all() and (oc[: int(w[0][-1])] == w[1]).all())
        cpu = {"qps": nl / t, "cores": CORES, "sample": nl}
    record(cfg, "locate()", npat, ms, best, 19 * 2 * 32 + 15.5 * 96 * tot / npat, par, cpu, {"occurrences": tot, "unit": "patterns/s"})
    # uniformly random 20-mers: almost all absent -> the search stops after ~4 symbols
    rflat = qr.integers(1, 256, npat * plen, dtype=np.uint8)
    d_rflat = dev(rflat)
    ms, best = time_gpu(lambda: csa.count(d_rflat, d_off), args.reps)
    record(cfg, "count() of random 20-mers (absent)", npat, ms, best, 4 * 2 * 32, None, None, {"unit": "patterns/s"})
    del d_rflat
    # the same index with the wavelet tree as its only occurrence structure (SDSLGPU_F_COMPACT)
    csa.close()
    torch.cuda.empty_cache()
    t0 = time.perf_counter()
    csa = pkg.CsaWt(text, flags=pkg.F_COMPACT)
    b2 = time.perf_counter() - t0
    ms, best = time_gpu(lambda: csa.count(d_flat, d_off), args.reps)
    par = bool((host(csa.count(d_flat, d_off)) == cnt).all())
    record(cfg, "count() [F_COMPACT: wt_huff only]", npat, ms, best, 7324, par, None, {"index_bytes": csa.device_bytes, "build_s": b2, "unit": "patterns/s"})
    ms, best = time_gpu(lambda: csa.locate(d_flat, d_off), args.reps)
    o2, c2 = csa.locate(d_flat, d_off)
    par = bool(torch.equal(o2, occ_off) and torch.equal(c2, occ))
    record(cfg, "locate() [F_COMPACT: wt_huff only]", npat, ms, best, 7324 + 2992 * tot / npat, par, None, {"occurrences": tot, "unit": "patterns/s"})
    del o2, c2
    if args.dense_sa:
        csa.close()
        torch.cuda.empty_cache()
        t0 = time.perf_counter()
        csa = pkg.CsaWt(text, sa_dens=args.dense_sa, isa_dens=args.dense_sa)
        b2 = time.perf_counter() - t0
        ms, best = time_gpu(lambda: csa.locate(d_flat, d_off), args.reps)
        o2, c2 = csa.locate(d_flat, d_off)
        par = bool(torch.equal(o2, occ_off) and torch.equal(c2, occ))
        record(cfg, f"locate() with t_dens = {args.dense_sa}", npat, ms, best, 19 * 2 * 32 + 96 * (args.dense_sa - 1) / 2 * tot / npat, par, None,
               {"occurrences": tot, "index_bytes": csa.device_bytes, "build_s": b2, "unit": "patterns/s"})
        del o2, c2
    if args.rrr_variant:
        csa.close()
        torch.cuda.empty_cache()
        t0 = time.perf_counter()
        csa = pkg.CsaWt(text, flags=pkg.F_RRR_BV)
        b2 = time.perf_counter() - t0
        ms, best = time_gpu(lambda: csa.count(d_flat, d_off), args.reps)
        par = bool((host(csa.count(d_flat, d_off)) == cnt).all())
        record(cfg, "count() on csa_wt<wt_huff<rrr_vector<63>>>", npat, ms, best, 7324, par, None, {"index_bytes": csa.device_bytes, "build_s": b2, "unit": "patterns/s"})
    csa.close()
    torch.cuda.empty_cache()


def main():
    global OUT
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="C2,C3,C4,C5")
    ap.add_argument("--nbits-log2", type=int, default=33)
    ap.add_argument("--queries", type=float, default=1e8)
    ap.add_argument("--queries-c3", type=float, default=1e8)
    ap.add_argument("--densities", default="0.01,0.05,0.1,0.25,0.5")
    ap.add_argument("--cpu-densities", default="0.1")
    ap.add_argument("--text-log2", type=int, default=28)
    ap.add_argument("--queries-c4", type=float, default=1e7)
    ap.add_argument("--csa-log2", type=int, default=28)
    ap.add_argument("--csa-ref", type=int, default=1)
    ap.add_argument("--rrr-variant", type=int, default=1)
    ap.add_argument("--dense-sa", type=int, default=4, help="also time locate() on an index with this t_dens (0 = skip)")
    ap.add_argument("--patterns", type=float, default=1e6)
    ap.add_argument("--cpu-sample", type=float, default=2e7)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    args.cpu_densities = [float(x) for x in args.cpu_densities.split(",") if x]
    if args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        OUT = open(args.out, "w")
    emit({"config": "env", "gpu": torch.cuda.get_device_name(0), "host_cores": CORES, "reference": REF is not None})
    for c in args.configs.split(","):
        t0 = time.perf_counter()
        {"C2": run_c2, "C3": run_c3, "C4": run_c4, "C5": run_c5}[c](args)
        emit({"config": c, "op": "wall_s", "seconds": time.perf_counter() - t0})


if __name__ == "__main__":
    main()
