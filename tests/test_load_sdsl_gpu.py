"""GPU: ingest of the reference's serialised bytes.  The blob comes from the unmodified
reference (oracle/_ref) where it is built, else from the oracle (whose serialisers are pinned byte-exact to it);
the loaded handle must answer exactly like a handle built from the raw input, and like the checker."""
import numpy as np
import pytest

import cases
import texts
from test_oracle_wt_int import queries as int_queries
from test_oracle_wt_int import sequences

pytestmark = pytest.mark.gpu


def _maker(oracle, orc):
    return orc.Ref() if orc.ref_available() else oracle


def test_load_bitvector_kinds(pkg, oracle, orc):
    mk = _maker(oracle, orc)
    for cid, w, nbits in cases.bitvector_catalogue(large=True):
        if nbits > 2_000_000:
            continue
        idx = cases.rank_queries(nbits, 3, 20000)
        blob = mk.bv(w, nbits).serialize(0)
        with pkg.load_sdsl(blob, pkg.KIND_BV) as v, pkg.BitVector(w, nbits) as direct:
            assert v.size == nbits and (v.rank(idx, 1) == direct.rank(idx, 1)).all(), cid
            q = cases.select_queries(v.arg_count(0), 4, 20000)
            if len(q):
                assert (v.select(q, 0) == direct.select(q, 0)).all(), cid
        for kind, name, cls in ((pkg.KIND_RRR63, "rrr", pkg.RrrVector), (pkg.KIND_SD, "sd", pkg.SdVector)):
            if name == "sd" and nbits == 0:
                continue
            chk = getattr(mk, name)(w, nbits)
            with pkg.load_sdsl(chk.serialize(), kind) as v:
                assert v.size == nbits
                for b in (0, 1):
                    assert (v.rank(idx, b) == chk.rank(idx, b)).all(), (cid, name, "rank", b)
                    q = cases.select_queries(v.arg_count(b), 5, 1500 if (name == "sd" and b == 0) else 20000)
                    if len(q):
                        assert (v.select(q, b) == chk.select(q, b)).all(), (cid, name, "select", b)
                pos = idx[idx < nbits]
                if len(pos):
                    assert (v.access(pos) == chk.access(pos)).all(), (cid, name, "access")
                if name == "rrr":  # the ingested image serialises back to the very same bytes
                    assert v.serialize() == chk.serialize(), (cid, "rrr round trip")


def test_load_wavelet_trees_and_csa(pkg, oracle, orc):
    mk = _maker(oracle, orc)
    rng = np.random.default_rng(5)
    for name, t in texts.text_catalogue(large=False):
        chk = mk.wt_huff(t)
        with pkg.load_sdsl(chk.serialize(), pkg.KIND_WT_HUFF) as wt:
            assert wt.size == len(t) and wt.sigma == len(set(t)), name
            i, c = texts.wt_queries(t, rng, 5000)
            assert (wt.rank(i, c) == chk.rank(i, c)).all(), (name, "rank")
            j = rng.integers(0, len(t), 5000, dtype=np.uint64)
            r, s = wt.inverse_select(j)
            rr, ss = chk.inverse_select(j)
            assert (r == rr).all() and (s == ss).all(), (name, "inverse_select")
            assert (wt.select(r + np.uint64(1), s.astype(np.uint8)) == j).all(), (name, "select round trip")
    for name, seq in sequences():
        chk = mk.wt_int(seq)
        with pkg.load_sdsl(chk.serialize(), pkg.KIND_WT_INT) as wt:
            i, c = int_queries(seq, rng, 5000)
            assert (wt.rank(i, c) == chk.rank(i, c)).all(), (name, "wt_int rank")
            j = rng.integers(0, len(seq), 5000, dtype=np.uint64)
            r, s = wt.inverse_select(j)
            assert (s == seq[j.astype(np.int64)]).all() and (wt.select(r + np.uint64(1), s) == j).all(), name
    for name, t in texts.text_catalogue(zero_free=True, large=False):
        chk = mk.csa(t)
        with pkg.load_sdsl(chk.serialize(), pkg.KIND_CSA_WT) as csa:
            assert csa.size == len(t) + 1
            pats = [t[s : s + int(rng.integers(1, 12))] for s in rng.integers(0, max(1, len(t) - 12), 400)] + [b"", b"\x01\x02zz"]
            flat, off = pkg.csr_patterns(pats)
            assert (csa.count(flat, off) == chk.count(flat, off)).all(), (name, "count")
            a, b = csa.locate(flat, off), chk.locate(flat, off)
            assert (a[0] == b[0]).all() and (a[1] == b[1]).all(), (name, "locate")
            k = rng.integers(0, len(t) + 1, 2000, dtype=np.uint64)
            assert (csa.sa(k) == chk.sa(k)).all(), (name, "SA access")


def test_load_rejects_garbage(pkg):
    for kind in (pkg.KIND_BV, pkg.KIND_RRR63, pkg.KIND_SD, pkg.KIND_WT_HUFF, pkg.KIND_WT_INT, pkg.KIND_CSA_WT):
        with pytest.raises(pkg.SdslGpuError):
            pkg.load_sdsl(b"\x01\x02\x03", kind)
    with pytest.raises(pkg.SdslGpuError):  # header promises more words than the blob holds
        pkg.load_sdsl((1 << 56 | 10_000).to_bytes(8, "little") + b"\0" * 16, pkg.KIND_BV)
