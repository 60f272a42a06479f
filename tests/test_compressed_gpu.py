"""GPU parity: rrr_vector<63> and sd_vector<> rank / select / access through the C ABI (SURVEY.md §8 rows a5,
a6), against the oracle and the unmodified reference; plus construction parity: the device encoders reproduce
the reference's serialised bytes."""
import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu

KINDS = {"rrr": "RrrVector", "sd": "SdVector"}


def _checkers(oracle, orc, kind, w, nbits):
    out = [("oracle", getattr(oracle, kind)(w, nbits))]
    if orc.ref_available():
        out.append(("reference", getattr(orc.Ref(), kind)(w, nbits)))
    return out


def _vectors():
    for cid, w, nbits in cases.bitvector_catalogue(large=True):
        yield cid, w, nbits
    for d in (0.01, 0.05, 0.25, 0.5, 0.9):
        nbits = (1 << 21) + 63 * 32 * 5
        yield f"density {d}", cases.bernoulli_words(nbits, d, int(d * 100)), nbits
    for nbits in (63, 63 * 32, 63 * 32 * 3, 63 * 31, 63 * 33, 2016 * 2 + 1):
        yield f"edge {nbits}", cases.random_words(nbits, nbits), nbits
        yield f"edge90 {nbits}", cases.bernoulli_words(nbits, 0.9, nbits), nbits


@pytest.mark.parametrize("kind", ["rrr", "sd"])
def test_compressed_catalogue(pkg, oracle, orc, kind):
    for cid, w, nbits in _vectors():
        if kind == "sd" and nbits == 0:
            continue
        with getattr(pkg, KINDS[kind])(w, nbits, flags=pkg.F_SDSL_LAYOUT) as v:
            assert v.size == nbits
            idx = cases.rank_queries(nbits, 11, 40000)
            got_rank = {b: v.rank(idx, b) for b in (0, 1)}
            sel_q = {b: cases.select_queries(v.arg_count(b), 12, 40000) for b in (0, 1)}
            got_sel = {b: v.select(sel_q[b], b) for b in (0, 1)}
            pos = idx[idx < nbits]
            got_acc = v.access(pos)
            blob = v.serialize()
            for name, chk in _checkers(oracle, orc, kind, w, nbits):
                for b in (0, 1):
                    assert (got_rank[b] == chk.rank(idx, b)).all(), (kind, cid, name, "rank", b)
                    assert v.arg_count(b) == int(chk.rank([nbits], b)[0]), (kind, cid, name, "arg_count")
                    if len(sel_q[b]):
                        assert (got_sel[b] == chk.select(sel_q[b], b)).all(), (kind, cid, name, "select", b)
                if len(pos):
                    assert (got_acc == chk.access(pos)).all(), (kind, cid, name, "access")
                ref_blob = chk.serialize()
                if kind == "rrr":
                    assert blob == ref_blob, (kind, cid, name, "serialised bytes")
                else:  # size, wl, low, high = a prefix of sd_vector::serialize
                    assert ref_blob[: len(blob)] == blob, (kind, cid, name, "serialised low/high")
            if kind == "rrr":  # beyond the last b-bit the reference answers size() in-band
                for b in (0, 1):
                    assert v.select(np.array([v.arg_count(b) + 1], np.uint64), b)[0] == nbits


@pytest.mark.parametrize("kind", ["rrr", "sd"])
def test_compressed_density_sweep_properties(pkg, oracle, kind):
    """BASELINE config 3 shape at 2^26 bits: density sweep, size-independent properties + oracle spot checks"""
    nbits = (1 << 26) + 4321
    for d in (0.01, 0.1, 0.5):
        w = cases.bernoulli_words(nbits, d, 7 + int(d * 1000))
        with getattr(pkg, KINDS[kind])(w, nbits) as v, pkg.BitVector(w, nbits) as plain:
            idx = cases.rank_queries(nbits, 8, 300000)
            r1 = v.rank(idx, 1)
            assert (r1 == plain.rank(idx, 1)).all() and (v.rank(idx, 0) + r1 == idx).all()
            k = cases.select_queries(v.arg_count(1), 9, 300000)
            p = v.select(k, 1)
            assert (p == plain.select(k, 1)).all() and (v.access(p) == 1).all()
            k0 = cases.select_queries(v.arg_count(0), 10, 300000)
            assert (v.select(k0, 0) == plain.select(k0, 0)).all()
            o = getattr(oracle, kind)(w, nbits)
            assert (v.rank(idx[:5000], 1) == o.rank(idx[:5000], 1)).all()


@pytest.mark.parametrize("kind", ["rrr", "sd"])
@pytest.mark.parametrize("chunk_bytes", ["256", "65536"])
def test_compressed_binned_order(pkg, oracle, monkeypatch, kind, chunk_bytes):
    """ORDER_BINNED (binned.cuh pipeline with the sd / rrr ops) forced onto the catalogue with tiny bins: rank (both
    patterns), select_1 and rrr select_0 agree with the oracle, out-of-domain queries included (rrr: the reference's
    in-band size() past the last b-bit); sd select_0 runs its own binned op over the sampled crossing blocks (sd_device.cuh)"""
    monkeypatch.setenv("SDSLGPU_BIN_CHUNK_BYTES", chunk_bytes)
    for cid, w, nbits in _vectors():
        if kind == "sd" and nbits == 0:
            continue
        o = getattr(oracle, kind)(w, nbits)
        with getattr(pkg, KINDS[kind])(w, nbits) as v:
            v.set_batch_order(pkg.ORDER_BINNED)
            for nq in (5, 8192, 20011):
                idx = cases.rank_queries(nbits, 3 + nq, nq)
                bad = np.arange(len(idx)) % 53 == 7
                idx_bad = idx.copy()
                idx_bad[bad] = np.uint64(nbits + 1 + nq)
                for b in (1, 0):
                    want = o.rank(idx, b)
                    assert (v.rank(idx, b) == want).all(), (kind, cid, "rank", b, nq)
                    want[bad] = pkg.NPOS
                    assert (v.rank(idx_bad, b) == want).all(), (kind, cid, "rank+ood", b, nq)
                m = v.arg_count(1)
                q = cases.select_queries(m, 4 + nq, nq)
                if len(q):
                    assert (v.select(q, 1) == o.select(q, 1)).all(), (kind, cid, "select1", nq)
                    qb = q.copy()
                    qb[::5] = 0
                    qb[2::5] = np.uint64(m + 1 + nq)
                    got = v.select(qb, 1)
                    assert (got[::5] == pkg.NPOS).all() and (got[1::5] == o.select(q[1::5], 1)).all(), (kind, cid, "select1 mixed", nq)
                    assert (got[2::5] == (nbits if kind == "rrr" else pkg.NPOS)).all(), (kind, cid, "select1 past the end", nq)
            q0 = cases.select_queries(v.arg_count(0), 6, 20000)
            if len(q0):
                assert (v.select(q0, 0) == o.select(q0, 0)).all(), (kind, cid, "select0")
                qb = q0.copy()
                qb[::7] = 0
                qb[3::7] = np.uint64(v.arg_count(0) + 1)
                got = v.select(qb, 0)
                assert (got[::7] == pkg.NPOS).all() and (got[1::7] == o.select(q0[1::7], 0)).all(), (kind, cid, "select0 mixed")
                assert (got[3::7] == (nbits if kind == "rrr" else pkg.NPOS)).all(), (kind, cid, "select0 past the end")
            if kind == "rrr":
                assert (v.select(np.array([0, v.arg_count(0) + 1, 2**63], np.uint64), 0) == np.array([pkg.NPOS, nbits, nbits], np.uint64)).all()
