#!/usr/bin/env python
"""Small inputs through every kernel added or rewritten in round 2, for compute-sanitizer (memcheck / racecheck):
chunk-position select samples, the int_vector<w> pack / unpack kernels, sd select_0 (sample table + crossing walk),
the rewritten rrr decode walks and adaptive hints, the wt_huff split builder, and a loopback group (fused peer-store
gather through the un-sort and the direct kernels, the copy kernel, the flag exchange).

    compute-sanitizer --tool memcheck python tools/sanitize_r02.py
    compute-sanitizer --tool racecheck python tools/sanitize_r02.py sort     (only the multi-tile sort / select-sector part)
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge  # noqa: E402
import cases  # noqa: E402
import texts  # noqa: E402

pkg, po = ge.load_package(), ge.load_oracle()
orc = po.Oracle()
rng = np.random.default_rng(3)


def dev(a):
    return torch.from_numpy(a.view(np.int64) if a.dtype == np.uint64 else a).cuda()


def host(t):
    a = t.cpu().numpy()
    return a.view(np.uint64) if a.dtype == np.int64 else a


# the persistent tile sort over several tiles per CTA (TMA copy of the next tile issued after the count pass, counters
# zeroed a phase early, three barriers per tile), ragged last tile, out-of-domain keys; select through select sectors
nbits = 300017
w = cases.bernoulli_words(nbits, 0.5, 5)
ob = orc.bv(w, nbits)
with pkg.BitVector(w, nbits) as bv:
    bv.set_batch_order(pkg.ORDER_BINNED)
    nq = 296 * 8192 * 2 + 4097  # more tiles than resident CTAs on any B200
    idx = rng.integers(0, nbits + 1, nq, dtype=np.uint64)
    idx[::1001] = np.uint64(nbits + 9)
    want = ob.rank(np.minimum(idx, np.uint64(nbits)), 1)
    want[::1001] = pkg.NPOS
    assert (host(bv.rank(dev(idx), 1)) == want).all()
    assert (host(bv.rank(dev(idx)[1:], 1)) == want[1:]).all()  # key array not 16-byte aligned: the plain-load sort
    for b in (1, 0):
        m = bv.arg_count(b)
        q = rng.integers(1, m + 1, nq, dtype=np.uint64)
        assert (host(bv.select(dev(q), b)) == ob.select(q, b)).all(), ("sectors", b)
if len(sys.argv) > 1 and sys.argv[1] == "sort":
    print("sanitize_r02 sort ok")
    sys.exit(0)

for nbits, dens in ((1, 0.5), (223, 0.5), (70001, 0.02), (300017, 0.5), (300017, 0.97)):
    w = cases.bernoulli_words(nbits, dens, 5)
    ob, osd, orr = orc.bv(w, nbits), orc.sd(w, nbits), orc.rrr(w, nbits)
    idx = rng.integers(0, nbits + 1, 5001, dtype=np.uint64)
    with pkg.BitVector(w, nbits) as bv, pkg.SdVector(w, nbits) as sd, pkg.RrrVector(w, nbits) as rrr:
        for order in (pkg.ORDER_DIRECT, pkg.ORDER_BINNED):
            for v, o in ((bv, ob), (sd, osd), (rrr, orr)):
                v.set_batch_order(order)
                for b in (1, 0):
                    assert (v.rank(idx, b) == o.rank(idx, b)).all()
                    m = v.arg_count(b)
                    if m:
                        q = rng.integers(1, m + 1, 5001, dtype=np.uint64)
                        assert (v.select(q, b) == o.select(q, b)).all(), (nbits, dens, order, b)
        for width in (20, 33):
            got = pkg.iv_unpack(bv.rank_iv(pkg.iv_pack(idx, width), width, len(idx), 1, width), width, len(idx))
            assert (got == ob.rank(idx, 1)).all()
            d = dev(pkg.iv_pack(idx, width))
            got = pkg.iv_unpack(host(bv.rank_iv(d, width, len(idx), 1, width)), width, len(idx))
            assert (got == ob.rank(idx, 1)).all()

for name, t in texts.text_catalogue(zero_free=True, large=False):
    if len(t) < 2:
        continue
    qi, qc = texts.wt_queries(t, rng, 2001)
    with pkg.WtHuff(t) as wt, pkg.WtHuff(t, flags=pkg.F_RRR_BV) as wtr:
        want = orc.wt_huff(t).rank(qi, qc)
        assert (wt.rank(qi, qc) == want).all() and (wtr.rank(qi, qc) == want).all(), name

with pkg.Group.create([0, 0, 0]) as g:
    nbits, nq = 200003, 30011
    w = cases.random_words(nbits, 9)
    ob = orc.bv(w, nbits)
    idx = rng.integers(0, nbits + 1, nq, dtype=np.uint64)
    hs = [pkg.BitVector(w, nbits) for _ in range(3)]
    sym = g.alloc(nq * 8)
    outs = [sym.tensor(k) for k in range(3)]
    d_idx = [dev(idx) for _ in range(3)]
    sel = rng.integers(1, hs[0].arg_count(1) + 1, nq, dtype=np.uint64)
    d_sel = [dev(sel) for _ in range(3)]
    for order in (pkg.ORDER_BINNED, pkg.ORDER_DIRECT):
        for h in hs:
            h.set_batch_order(order)
        for gm in (pkg.GATHER_FUSED, pkg.GATHER_PACKED):
            g.rank(hs, 1, d_idx, outs, gather=gm)
            for o in outs:
                assert (host(o) == ob.rank(idx, 1)).all()
            g.select(hs, 1, d_sel, outs, gather=gm)
            for o in outs:
                assert (host(o) == ob.select(sel, 1)).all()
    t = dict(texts.text_catalogue(zero_free=True, large=False))["dna"]
    qi, qc = texts.wt_queries(t, rng, nq)
    wts = [pkg.WtHuff(t) for _ in range(3)]
    for gm in (pkg.GATHER_FUSED, pkg.GATHER_PACKED):
        g.wt_rank(wts, [dev(qi)] * 3, [dev(qc)] * 3, outs, gather=gm)
        for o in outs:
            assert (host(o) == orc.wt_huff(t).rank(qi, qc)).all()
    sym.release()
    for h in hs + wts:
        h.close()
print("sanitize_r02 ok")
