"""GPU: the multi-GPU data plane of the C ABI (sdslgpu_group_*, csrc/group.cu) against the oracle.

 * loopback groups (several members on ONE device) run on any GPU box: they exercise sharding, the fused peer-store
   gather (bin_unsort_kernel<true>, fan_copy_kernel) and the flag exchange without a second GPU;
 * real multi-device groups (one process driving N devices; one process per GPU under torchrun) need >= 2 GPUs and
   are skipped otherwise: NCCL all-gather, fused gather over NVLink, replicate."""
import os
import subprocess
import sys

import numpy as np
import pytest

import cases
import texts

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch

    return torch.cuda.device_count()


def _dev(a, d):
    import torch

    t = torch.from_numpy(a.view(np.int64) if a.dtype == np.uint64 else a)
    return t.to(torch.device("cuda", d))


def _host(t):
    a = t.cpu().numpy()
    return a.view(np.uint64) if a.dtype == np.int64 else a


def _run_bv(pkg, oracle, g, devices, gathers, nbits=3_000_001, nq=100_003, orders=None):
    import torch

    w = cases.random_words(nbits, 11)
    ob = oracle.bv(w, nbits)
    rng = np.random.default_rng(5)
    idx = rng.integers(0, nbits + 1, nq, dtype=np.uint64)
    idx[:3] = [0, nbits, nbits + 5]  # the last one is out of domain -> NPOS
    want_r = {b: ob.rank(np.minimum(idx, np.uint64(nbits)), b) for b in (0, 1)}
    hs = [pkg.BitVector(w, nbits, device=d) for d in devices]
    try:
        m1 = hs[0].arg_count(1)
        sel = rng.integers(1, m1 + 1, nq, dtype=np.uint64)
        want_s = ob.select(sel, 1)
        sel[-1] = 0  # out of domain
        d_idx = [_dev(idx, d) for d in devices]
        d_sel = [_dev(sel, d) for d in devices]
        sym = g.alloc(2 * nq * 8)
        outs = [sym.tensor(k)[:nq] for k in range(len(devices))]
        outs2 = [sym.tensor(k)[nq:] for k in range(len(devices))]
        plain = [torch.empty(nq, dtype=torch.int64, device=torch.device("cuda", d)) for d in devices]
        for order in orders or (pkg.ORDER_BINNED, pkg.ORDER_DIRECT):
            for h in hs:
                h.set_batch_order(order)
            for gather in gathers:
                for b in (1, 0):
                    for o in outs:
                        o.fill_(-7)
                    g.rank(hs, b, d_idx, outs, gather=gather)
                    for k, o in enumerate(outs):
                        got = _host(o)
                        assert (got[:2] == want_r[b][:2]).all() and got[2] == pkg.NPOS
                        assert (got[3:] == want_r[b][3:]).all(), ("rank", order, gather, b, k)
                for o in outs2:
                    o.fill_(-7)
                g.select(hs, 1, d_sel, outs2, gather=gather)
                for k, o in enumerate(outs2):
                    got = _host(o)
                    assert (got[:-1] == want_s[:-1]).all() and got[-1] == pkg.NPOS, ("select", order, gather, k)
            for gm in (pkg.GATHER_NCCL, pkg.GATHER_PACKED):  # both work on any device memory, not only group-allocated
                if gm in gathers:
                    for o in plain:
                        o.fill_(-7)
                    g.rank(hs, 1, d_idx, plain, gather=gm)
                    for o in plain:
                        got = _host(o)
                        assert (got[3:] == want_r[1][3:]).all() and got[2] == pkg.NPOS, gm
        # GATHER_NONE: every member holds its own shard (and the left-over tail)
        for o in outs:
            o.fill_(-7)
        g.rank(hs, 1, d_idx, outs, gather=pkg.GATHER_NONE)
        s = nq // g.nranks
        for k, o in enumerate(outs):
            got = _host(o)
            r = g.first_rank + k
            lo, hi = r * s, (r + 1) * s
            keep = np.ones(nq, bool)
            keep[:3] = False
            sl = np.zeros(nq, bool)
            sl[lo:hi] = True
            sl[s * g.nranks:] = True
            assert (got[sl & keep] == want_r[1][sl & keep]).all()
            if g.nranks > 1 and g.nlocal == g.nranks:
                assert (got[~sl] == np.uint64(2**64 - 7)).all()
        sym.release()
    finally:
        for h in hs:
            h.close()


def _run_compressed(pkg, oracle, g, devices, gathers):
    """rrr_vector<63> / sd_vector<> through the same group calls (copy-kernel gather)"""
    nbits, nq = 1_000_003, 40_001
    w = cases.bernoulli_words(nbits, 0.2, 21)
    rng = np.random.default_rng(6)
    idx = rng.integers(0, nbits + 1, nq, dtype=np.uint64)
    for kind, cls in (("rrr", pkg.RrrVector), ("sd", pkg.SdVector)):
        o = getattr(oracle, kind)(w, nbits)
        hs = [cls(w, nbits, device=d) for d in devices]
        try:
            sym = g.alloc(nq * 8)
            outs = [sym.tensor(k) for k in range(len(devices))]
            d_idx = [_dev(idx, d) for d in devices]
            for gather in gathers:
                for b in (1, 0):
                    for t in outs:
                        t.fill_(-7)
                    g.rank(hs, b, d_idx, outs, gather=gather)
                    for t in outs:
                        assert (_host(t) == o.rank(idx, b)).all(), (kind, "rank", b, gather)
                    sel = rng.integers(1, hs[0].arg_count(b) + 1, nq, dtype=np.uint64)
                    g.select(hs, b, [_dev(sel, d) for d in devices], outs, gather=gather)
                    for t in outs:
                        assert (_host(t) == o.select(sel, b)).all(), (kind, "select", b, gather)
            sym.release()
        finally:
            for h in hs:
                h.close()


def _run_wt_fm(pkg, oracle, g, devices, gathers):
    rng = np.random.default_rng(9)
    t = dict(texts.text_catalogue(zero_free=True, large=False))["dna"]
    qi, qc = texts.wt_queries(t, rng, 50_001)
    want = oracle.wt_huff(t).rank(qi, qc)
    wts = [pkg.WtHuff(t, device=d) for d in devices]
    csas = [pkg.CsaWt(t, device=d) for d in devices]
    try:
        nq = len(qi)
        sym = g.alloc(nq * 8)
        outs = [sym.tensor(k) for k in range(len(devices))]
        d_i = [_dev(qi, d) for d in devices]
        d_c = [_dev(qc, d) for d in devices]
        for gather in gathers:
            for o in outs:
                o.fill_(-7)
            g.wt_rank(wts, d_i, d_c, outs, gather=gather)
            for o in outs:
                assert (_host(o) == want).all(), ("wt_rank", gather)
        pats = [t[s : s + 7] for s in rng.integers(0, len(t) - 7, 20_011)] + [b"", b"zzzz"]
        flat, off = pkg.csr_patterns(pats)
        wantc = oracle.csa(t).count(flat, off)
        d_f = [_dev(flat, d) for d in devices]
        d_o = [_dev(off, d) for d in devices]
        np_ = len(pats)
        for gather in gathers:
            for o in outs:
                o.fill_(-7)
            g.fm_count(csas, d_f, d_o, [o[:np_] for o in outs], gather=gather)
            for o in outs:
                assert (_host(o[:np_]) == wantc).all(), ("fm_count", gather)
        sym.release()
    finally:
        for h in wts + csas:
            h.close()


@pytest.mark.parametrize("members", [2, 3])
def test_group_loopback_fused_on_one_gpu(pkg, oracle, members):
    devices = [0] * members
    with pkg.Group.create(devices) as g:
        assert g.nranks == members and g.nlocal == members and g.fused_possible
        _run_bv(pkg, oracle, g, devices, [pkg.GATHER_FUSED, pkg.GATHER_PACKED, pkg.GATHER_AUTO])
        _run_bv(pkg, oracle, g, devices, [pkg.GATHER_PACKED], nbits=1 << 24, nq=70_001)  # 26-bit fields
        _run_compressed(pkg, oracle, g, devices, [pkg.GATHER_FUSED, pkg.GATHER_PACKED])
        _run_wt_fm(pkg, oracle, g, devices, [pkg.GATHER_FUSED, pkg.GATHER_PACKED])
        import torch

        with pytest.raises(pkg.SdslGpuError):  # FUSED needs group-allocated result arrays
            h = pkg.BitVector(np.zeros(4, np.uint64), 200)
            q = [torch.zeros(10, dtype=torch.int64, device="cuda")] * members
            o = [torch.zeros(10, dtype=torch.int64, device="cuda") for _ in range(members)]
            try:
                g.rank([h] * members, 1, q, o, gather=pkg.GATHER_FUSED)
            finally:
                h.close()


def test_group_of_one(pkg, oracle):
    with pkg.Group.create([0]) as g:
        assert g.nranks == 1
        _run_bv(pkg, oracle, g, [0], [pkg.GATHER_AUTO, pkg.GATHER_NCCL, pkg.GATHER_FUSED], nq=20_001, orders=[pkg.ORDER_AUTO])


def test_group_multi_device_nccl_and_fused(pkg, oracle):
    n = _ngpu()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    devices = list(range(min(n, 8)))
    with pkg.Group.create(devices) as g:
        assert g.nranks == len(devices)
        gathers = [pkg.GATHER_NCCL] + ([pkg.GATHER_FUSED, pkg.GATHER_PACKED, pkg.GATHER_AUTO] if g.fused_possible else [])
        _run_bv(pkg, oracle, g, devices, gathers)
        _run_compressed(pkg, oracle, g, devices, gathers)
        _run_wt_fm(pkg, oracle, g, devices, gathers)
        # replicate: an index built on device 0 arrives on every member and answers identically
        nbits = 1_000_003
        w = cases.random_words(nbits, 3)
        idx = np.random.default_rng(1).integers(0, nbits + 1, 30_000, dtype=np.uint64)
        with pkg.BitVector(w, nbits, device=0) as src:
            reps = g.replicate(src, root=0)
            want = oracle.bv(w, nbits).rank(idx, 1)
            for d, r in zip(devices, reps):
                assert (_host(r.rank(_dev(idx, d), 1)) == want).all()
                r.close()
        t = dict(texts.text_catalogue(zero_free=True, large=False))["dna"]
        with pkg.CsaWt(t, device=0) as src:
            reps = g.replicate(src, root=0)
            pats = [t[s : s + 5] for s in range(0, 4000, 7)]
            flat, off = pkg.csr_patterns(pats)
            want = oracle.csa(t).count(flat, off)
            for d, r in zip(devices, reps):
                assert (_host(r.count(_dev(flat, d), _dev(off, d))) == want).all()
                r.close()


def test_group_one_process_per_gpu_torchrun(pkg):
    n = _ngpu()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = min(n, 4)
    port = 29600 + os.getpid() % 300
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "mp_group_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0 and r.stdout.count("group worker ok") == world, (r.stdout[-3000:], r.stderr[-3000:])
