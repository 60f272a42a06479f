#!/usr/bin/env python
"""small forced-binned rank/select batches for compute-sanitizer (run with SDSLGPU_BIN_CHUNK_BYTES=4096)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

pkg, orc = ge.load_package(), ge.load_oracle()
rng = np.random.default_rng(1)
for nbits in (0, 1, 223, 100001, 3_000_017):
    w = rng.integers(0, 2**64, (nbits + 63) // 64, dtype=np.uint64)
    o = orc.Oracle().bv(w, nbits)
    with pkg.BitVector(w, nbits) as bv:
        bv.set_batch_order(pkg.ORDER_BINNED)
        for nq in (1, 8191, 8192, 30011):
            idx = rng.integers(0, nbits + 3, nq, dtype=np.uint64)
            for b in (1, 0):
                want = o.rank(np.minimum(idx, nbits), b)
                want[idx > nbits] = pkg.NPOS
                assert (bv.rank(idx, b) == want).all(), ("rank", nbits, nq, b)
                m = bv.arg_count(b)
                q = rng.integers(0, m + 2, nq, dtype=np.uint64)
                ok = (q >= 1) & (q <= m)
                got = bv.select(q, b)
                assert (got[~ok] == pkg.NPOS).all()
                if ok.any():
                    assert (got[ok] == o.select(q[ok], b)).all(), ("select", nbits, nq, b)
print("sanitize_binned ok")
