"""GPU parity: wt_int<> rank / select / access / inverse_select through the C ABI
against the oracle and the unmodified reference (the shape of test/wt_int_test.cpp:125-190)."""
import numpy as np
import pytest

from test_oracle_wt_int import queries, sequences

pytestmark = pytest.mark.gpu


def test_wt_int_catalogue(pkg, oracle, orc):
    rng = np.random.default_rng(10)
    for name, seq in sequences():
        n = len(seq)
        with pkg.WtInt(seq) as wt:
            assert wt.size == n and wt.sigma == len(np.unique(seq)), name
            i, c = queries(seq, rng, 20000)
            got_rank = wt.rank(i, c)
            tot = wt.rank(np.full(len(c), n, dtype=np.uint64), c)
            ok = tot > 0
            k = (rng.integers(0, 2**62, len(c), dtype=np.uint64) % np.maximum(tot, 1)) + np.uint64(1)
            got_sel = wt.select(k[ok], c[ok])
            j = rng.integers(0, n, 20000, dtype=np.uint64)
            got_rnk, got_sym = wt.inverse_select(j)
            assert (got_sym == seq[j.astype(np.int64)]).all() and (wt.access(j) == got_sym).all(), name
            checkers = [("oracle", oracle.wt_int(seq))]
            if orc.ref_available():
                checkers.append(("reference", orc.Ref().wt_int(seq)))
            for cname, chk in checkers:
                assert (got_rank == chk.rank(i, c)).all(), (name, cname, "rank")
                assert (got_sel == chk.select(k[ok], c[ok])).all(), (name, cname, "select")
                rr, ss = chk.inverse_select(j)
                assert (got_rnk == rr).all() and (got_sym == ss).all(), (name, cname, "inverse_select")
            # defined results where the reference throws / is undefined
            assert (wt.select(tot[ok] + np.uint64(1), c[ok]) == pkg.NPOS).all()
            assert (wt.select(got_rnk + np.uint64(1), got_sym) == j).all(), (name, "round trip")


def test_wt_int_large_properties(pkg):
    rng = np.random.default_rng(11)
    n = 1 << 24
    seq = rng.integers(0, 1 << 20, n, dtype=np.uint64)
    with pkg.WtInt(seq) as wt:
        j = rng.integers(0, n, 500000, dtype=np.uint64)
        rnk, sym = wt.inverse_select(j)
        assert (sym == seq[j.astype(np.int64)]).all()
        assert (wt.select(rnk + np.uint64(1), sym) == j).all()
        assert (wt.rank(j + np.uint64(1), sym) == rnk + np.uint64(1)).all()
