"""CPU: the C restatement of wt_int<> (oracle/oracle_wt.c) against the UNMODIFIED reference — row a8:
serialised bytes, rank / select / inverse_select (the shape of test/wt_int_test.cpp:125-190)."""
import numpy as np


def sequences():
    rng = np.random.default_rng(8)
    yield "tiny", np.array([3, 1, 4, 1, 5, 9, 2, 6, 5, 3, 5], dtype=np.uint64)
    yield "zeros", np.zeros(100, dtype=np.uint64)
    yield "ones", np.ones(77, dtype=np.uint64)
    yield "single", np.array([12345], dtype=np.uint64)
    yield "bytes", rng.integers(0, 256, 20000, dtype=np.uint64)
    yield "wide", rng.integers(0, 1 << 40, 5000, dtype=np.uint64)
    yield "skewed", np.minimum(rng.geometric(0.01, 30000), 5000).astype(np.uint64)
    yield "sorted", np.sort(rng.integers(0, 100000, 10000, dtype=np.uint64))
    yield "pow2", (np.uint64(1) << rng.integers(0, 20, 3000, dtype=np.uint64))


def queries(seq, rng, nq):
    n = len(seq)
    i = rng.integers(0, n + 1, nq, dtype=np.uint64)
    c = seq[rng.integers(0, n, nq)].copy()
    c[::3] = rng.integers(0, int(seq.max()) * 2 + 3, len(c[::3]), dtype=np.uint64)
    return i, c


def test_wt_int_vs_reference(oracle, ref):
    rng = np.random.default_rng(9)
    for name, seq in sequences():
        ow, rw = oracle.wt_int(seq), ref.wt_int(seq)
        assert ow.serialize() == rw.serialize(), (name, "serialised bytes")
        n = len(seq)
        i, c = queries(seq, rng, 5000)
        assert (ow.rank(i, c) == rw.rank(i, c)).all(), (name, "rank")
        tot = rw.rank(np.full(len(c), n, dtype=np.uint64), c)
        ok = tot > 0
        k = (rng.integers(0, 2**62, len(c), dtype=np.uint64) % np.maximum(tot, 1)) + np.uint64(1)
        assert (ow.select(k[ok], c[ok]) == rw.select(k[ok], c[ok])).all(), (name, "select")
        j = rng.integers(0, n, 5000, dtype=np.uint64)
        a, b = ow.inverse_select(j), rw.inverse_select(j)
        assert (a[0] == b[0]).all() and (a[1] == b[1]).all() and (a[1] == seq[j.astype(np.int64)]).all(), (name, "inverse_select")
