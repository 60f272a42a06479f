"""tools/bench_extras.py — the rest of BASELINE.json's metric, measured by bench.py next to the config[1] headline:

    fm_count   C5  csa_wt<wt_huff<>> on a 2^30-byte text, 1e6 count() of |P| = 20          (patterns/s)
    wt_rank    C4  wt_huff<> on 2^28 uniform bytes (sigma 256), 1e7 rank(i, c)            (queries/s)
    rrr / sd   C3  rrr_vector<63> / sd_vector<> on 2^33 bits at one density, 1e8 rank_1 + 1e8 select_1

Each function returns one record: device-resident CUDA-event time (median of `reps` after 2 warm-ups), the rate, a
`roofline` block (algorithmic bytes of SURVEY.md §8(d) against the measured HBM peak), a `parity` block (a sample of the
timed results compared with the UNMODIFIED reference, oracle/_ref) and a `cpu_baseline` (the reference's rate on the
host cores for the same sample).  The oracle / reference are used here as CHECKER and BASELINE only.
"""
import os
import time

import numpy as np
import torch

CORES = os.cpu_count() or 1


def dev(a, device=None):
    t = torch.from_numpy(a.view(np.int64) if a.dtype == np.uint64 else a)
    return t.cuda() if device is None else t.to(device)


def host(t):
    a = t.cpu().numpy()
    return a.view(np.uint64) if a.dtype == np.int64 else a


def time_gpu(fn, reps=5, warm=2):
    for _ in range(warm):
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
    return ts[len(ts) // 2]


def roofline(units, bytes_per_unit, ms, peak, peak_src, read_bytes_per_unit=None, note=None):
    a = units * bytes_per_unit / (ms * 1e-3) / 1e9
    r = {"bound": "hbm", "achieved": a, "peak": peak, "unit": "GB/s", "frac": a / peak, "traffic": None, "peak_source": peak_src,
         "algorithmic_bytes_per_unit": bytes_per_unit, "kernel_ms": ms}
    if read_bytes_per_unit is not None:
        r["frac_read_only"] = units * read_bytes_per_unit / (ms * 1e-3) / 1e9 / peak
        r["read_bytes_per_unit"] = read_bytes_per_unit
    if note:
        r["note"] = note
    return r


def gpu_random_words(nbits, density, seed):
    """Bernoulli(density) bit vector generated on the device -> int64 words tensor"""
    g = torch.Generator(device="cuda").manual_seed(seed)
    nw = (nbits + 63) // 64
    out = torch.empty(nw, dtype=torch.int64, device="cuda")
    chunk = 1 << 21
    w = torch.ones(64, dtype=torch.int64, device="cuda") << torch.arange(64, device="cuda", dtype=torch.int64)
    for lo in range(0, nw, chunk):
        hi = min(nw, lo + chunk)
        bits = (torch.rand((hi - lo, 64), device="cuda", generator=g) < density).to(torch.int64)
        out[lo:hi] = (bits * w).sum(1)
    return out


def c5_workload(log2n, npat, plen=20, seed=42):
    rng = np.random.default_rng(seed)
    n = 1 << log2n
    text = rng.integers(1, 256, n, dtype=np.uint8)
    starts = np.random.default_rng(7).integers(0, n - plen, npat)
    flat = text[(starts[:, None] + np.arange(plen)[None, :])].reshape(-1).copy()
    off = np.arange(npat + 1, dtype=np.uint64) * np.uint64(plen)
    return text, flat, off


def fm_count_record(pkg, po, peak, peak_src, log2n=30, npat=1_000_000, reps=5, sample=20000, with_reference=True, csa=None, workload=None):
    """C5.  The default index holds the 32 one-hot occurrence bitmaps (5.6 B/symbol): 2 gathers per
    backward-search step; the SDSL-basis roofline charges the reference's 7324 B per pattern (19 steps x 2 ends x ~8
    levels x 24 B), the occ16 basis the 1216 B this structure really needs."""
    text, flat, off = workload or c5_workload(log2n, npat)
    t0 = time.perf_counter()
    own = csa is None
    if own:
        csa = pkg.CsaWt(text)
    build_s = time.perf_counter() - t0
    d_flat, d_off = dev(flat), dev(off)
    ms = time_gpu(lambda: csa.count(d_flat, d_off), reps)
    cnt = host(csa.count(d_flat, d_off))
    rec = {"config": f"C5 csa_wt<wt_huff<>> on a 2^{log2n}-byte uniform text (bytes 1..255), {npat} count() of |P|=20 sampled from the text",
           "value": npat / (ms * 1e-3), "unit": "patterns/s", "ms": ms, "index_device_bytes": csa.device_bytes, "gpu_build_s": build_s,
           "index_note": "default index = wt_huff of the BWT + 32 one-hot sector-block bitmaps + the BWT (5.6 B/symbol; the reference's csa_wt is ~1.9 B/symbol); "
                         "SDSLGPU_F_COMPACT keeps the tree only (identical results, ~3x slower count)",
           "roofline": roofline(npat, 1216, ms, peak, peak_src, note="occ16 basis: 19 steps x 2 gathers x 32 B per pattern"),
           "roofline_sdsl_basis": roofline(npat, 7324, ms, peak, peak_src, note="SDSL basis (SURVEY §8(d)): what the reference's wt_huff cascade would read"),
           "all_found": bool((cnt >= 1).all())}
    parity, cpu = None, None
    if with_reference and po.ref_available():
        t0 = time.perf_counter()
        blob = csa.serialize(0)  # the reference's own byte format, written by this engine
        egress_s = time.perf_counter() - t0
        t0 = time.perf_counter()
        ref = po.Ref().csa(blob=blob)
        load_s = time.perf_counter() - t0
        del blob
        ns = min(npat, sample)
        t0 = time.perf_counter()
        want = ref.count(flat[: ns * 20], off[: ns + 1], threads=CORES)
        t = time.perf_counter() - t0
        parity = {"checked_patterns": ns, "against": "reference (loaded the index this engine serialised)", "bit_exact": bool((cnt[:ns] == want).all())}
        cpu = {"value": ns / t, "unit": "patterns/s", "cores": CORES, "kind": "reference", "sample": f"{ns} of the timed patterns on the same index",
               "egress_s": egress_s, "reference_load_s": load_s}
        del ref
    rec["parity"], rec["cpu_baseline"] = parity, cpu
    if own:
        csa.close()
    return rec


def wt_rank_record(pkg, po, peak, peak_src, log2n=28, nq=10_000_000, reps=5, sample=2_000_000, with_reference=True):
    """C4: level-synchronous wt_huff rank; 209 B per query on the reference's layout (8 levels x 24 B + query/result)"""
    n = 1 << log2n
    text = np.random.default_rng(42).integers(0, 256, n, dtype=np.uint8)
    t0 = time.perf_counter()
    wt = pkg.WtHuff(text)
    build_s = time.perf_counter() - t0
    qr = np.random.default_rng(7)
    i = qr.integers(0, n + 1, nq, dtype=np.uint64)
    c = qr.integers(0, 256, nq, dtype=np.uint8)
    d_i, d_c, d_out = dev(i), dev(c), torch.empty(nq, dtype=torch.int64, device="cuda")
    ms = time_gpu(lambda: wt.rank(d_i, d_c, out=d_out), reps)
    rec = {"config": f"C4 wt_huff<> on 2^{log2n} uniform bytes (sigma 256), {nq} rank(i,c)", "value": nq / (ms * 1e-3), "unit": "queries/s", "ms": ms,
           "index_device_bytes": wt.device_bytes, "gpu_build_s": build_s, "roofline": roofline(nq, 209, ms, peak, peak_src, read_bytes_per_unit=201)}
    parity, cpu = None, None
    if with_reference and po.ref_available():
        ns = min(nq, sample)
        t0 = time.perf_counter()
        ref = po.Ref().wt_huff(text)
        rb = time.perf_counter() - t0
        t0 = time.perf_counter()
        want = ref.rank(i[:ns], c[:ns], threads=CORES)
        t = time.perf_counter() - t0
        parity = {"checked_queries": ns, "against": "reference", "bit_exact": bool((host(d_out[:ns]) == want).all())}
        cpu = {"value": ns / t, "unit": "queries/s", "cores": CORES, "kind": "reference", "sample": f"{ns} of the timed queries", "reference_build_s": rb}
        del ref
    rec["parity"], rec["cpu_baseline"] = parity, cpu
    wt.close()
    return rec


def compressed_records(pkg, po, peak, peak_src, log2n=33, density=0.1, nq=100_000_000, reps=3, sample=2_000_000, with_reference=True):
    """C3 at one density: rrr_vector<63> and sd_vector<> rank_1 / select_1 (+ sd select_0)"""
    nbits = 1 << log2n
    words_d = gpu_random_words(nbits, density, 42 + int(density * 100))
    qr = np.random.default_rng(7)
    idx = qr.integers(0, nbits + 1, nq, dtype=np.uint64)
    d_idx = dev(idx)
    d_out = torch.empty(nq, dtype=torch.int64, device="cuda")
    words_h = host(words_d) if (with_reference and po.ref_available()) else None
    out = {}
    for name, cls, rank_b, sel_b in (("rrr", pkg.RrrVector, 76, 76), ("sd", pkg.SdVector, 72, 56)):
        t0 = time.perf_counter()
        v = cls(words_d, nbits)
        torch.cuda.synchronize()
        build_s = time.perf_counter() - t0
        m = v.arg_count(1)
        sel = qr.integers(1, m + 1, nq, dtype=np.uint64)
        d_sel = dev(sel)
        ref = None
        if words_h is not None:
            t0 = time.perf_counter()
            ref = (po.Ref().rrr if name == "rrr" else po.Ref().sd)(words_h, nbits)
            rb = time.perf_counter() - t0
        ops = {}
        for op, q, dq, nbytes in (("rank_1", idx, d_idx, rank_b), ("select_1", sel, d_sel, sel_b)):
            fn = (lambda: v.rank(dq, 1, out=d_out)) if op == "rank_1" else (lambda: v.select(dq, 1, out=d_out))
            ms = time_gpu(fn, reps)
            o = {"value": nq / (ms * 1e-3), "unit": "queries/s", "ms": ms, "roofline": roofline(nq, nbytes, ms, peak, peak_src, read_bytes_per_unit=nbytes - 8)}
            if ref is not None:
                ns = min(nq, sample)
                t0 = time.perf_counter()
                want = ref.rank(q[:ns], 1, threads=CORES) if op == "rank_1" else ref.select(q[:ns], 1, threads=CORES)
                t = time.perf_counter() - t0
                o["parity"] = {"checked_queries": ns, "against": "reference", "bit_exact": bool((host(d_out[:ns]) == want).all())}
                o["cpu_baseline"] = {"value": ns / t, "unit": "queries/s", "cores": CORES, "kind": "reference", "sample": f"{ns} of the timed queries",
                                     "reference_build_s": rb}
            ops[op] = o
        out[name] = {"config": f"C3 {'rrr_vector<63>' if name == 'rrr' else 'sd_vector<>'} on 2^{log2n} bits, density {density:g}, {nq} queries per op",
                     "index_device_bytes": v.device_bytes, "bits_per_bit": 8.0 * v.device_bytes / nbits, "gpu_build_s": build_s, **ops}
        v.close()
        del ref, d_sel
        torch.cuda.empty_cache()
    return out
