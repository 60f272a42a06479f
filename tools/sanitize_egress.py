#!/usr/bin/env python
"""small inputs through the device wavelet-tree builders (wt_build.cu) and the egress path (sdsl_egress.cu) for
compute-sanitizer; every blob is compared with the checker's (reference where built, else the oracle)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

pkg, orc = ge.load_package(), ge.load_oracle()
mk = orc.Ref() if orc.ref_available() else orc.Oracle()
rng = np.random.default_rng(3)
deep = np.concatenate([np.full(1 << k, 65 + k, np.uint8) for k in range(12)])
rng.shuffle(deep)
for name, t in (("one", b"\n"), ("aaa", b"a" * 100), ("abc", b"abc_abc_abc\n"), ("deep", deep.tobytes()),
                ("uniform", rng.integers(0, 256, 40000, dtype=np.uint8).tobytes())):
    with pkg.WtHuff(t) as wt:
        assert wt.serialize() == mk.wt_huff(t).serialize(), name
    if 0 not in t:
        with pkg.CsaWt(t) as csa:
            assert csa.serialize() == mk.csa(t).serialize(), name
for seq in (np.array([5], np.uint64), np.zeros(70, np.uint64), rng.integers(0, 1 << 33, 3000, dtype=np.uint64), rng.integers(0, 7, 5000, dtype=np.uint64)):
    with pkg.WtInt(seq) as wi:
        assert wi.serialize() == mk.wt_int(seq).serialize()
for nbits, dens in ((0, 0.5), (1, 1.0), (4097, 0.5), (100001, 0.5), (300000, 0.001), (300000, 0.04)):
    bits = (rng.random(nbits) < dens).astype(np.uint8)
    w = np.packbits(np.concatenate([bits, np.zeros((-nbits) % 64, np.uint8)]), bitorder="little").view(np.uint64) if nbits else np.zeros(0, np.uint64)
    chk = mk.bv(w, nbits)
    with pkg.BitVector(w, nbits) as bv:
        for what in range(5):
            assert bv.serialize(what) == chk.serialize(what), (nbits, dens, what)
    if nbits > 1:
        with pkg.SdVector(w, nbits) as sd:
            assert sd.serialize(1) == mk.sd(w, nbits).serialize(), (nbits, dens)
print("sanitize_egress ok")
