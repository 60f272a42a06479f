"""GPU: egress (SURVEY.md §8(f)-1 "... and back"): sdslgpu_serialize must return, byte for byte, what the unmodified
reference's serialize() / store_to_file writes for the same input — select_support_mcl<1>/<0>, sd_vector<>,
wt_huff<>, wt_int<>, csa_wt<wt_huff<>> — and the reference must be able to load and query the blob.  The checker is
oracle/_ref where it was built, else the oracle (its serialisers are pinned byte-exact to the reference on CPU)."""
import numpy as np
import pytest

import cases
import texts
from test_oracle_wt_int import sequences

pytestmark = pytest.mark.gpu


def _maker(oracle, orc):
    return orc.Ref() if orc.ref_available() else oracle


def _clean(w, nbits):
    w = np.array(w, dtype=np.uint64, copy=True)
    if nbits % 64:
        w[-1] &= np.uint64((1 << (nbits % 64)) - 1)  # bits past size() are unspecified in the reference; the library drops them
    return w


def test_select_supports_and_sd_vector(pkg, oracle, orc):
    mk = _maker(oracle, orc)
    for cid, w, nbits in cases.bitvector_catalogue(large=True):
        w = _clean(w, nbits)
        chk = mk.bv(w, nbits)
        with pkg.BitVector(w, nbits) as v:  # default layout: sector blocks only, nothing of the reference's form resident
            assert v.serialize(0) == chk.serialize(0), (cid, "bit_vector")
            assert v.serialize(1) == chk.serialize(1), (cid, "rank_support_v<1>")
            assert v.serialize(2) == chk.serialize(2), (cid, "rank_support_v<0>")
            assert v.serialize(3) == chk.serialize(3), (cid, "select_support_mcl<1>")
            assert v.serialize(4) == chk.serialize(4), (cid, "select_support_mcl<0>")
            assert v.serialize(5) == chk.serialize(5), (cid, "rank_support_v5<1>")
            assert v.serialize(6) == chk.serialize(6), (cid, "rank_support_v5<0>")
        if nbits and nbits <= 2_000_000:
            with pkg.SdVector(w, nbits) as v:
                assert v.serialize(1) == mk.sd(w, nbits).serialize(), (cid, "sd_vector")


def test_select_support_block_edges(pkg, oracle, orc):
    """4032 / 4033 arguments in the trailing superblock, exact multiples of 4096, both construction modes"""
    mk = _maker(oracle, orc)
    rng = np.random.default_rng(12)
    for n in (99999, 100000, 300000):
        for m in (1, 64, 65, 4032, 4033, 4095, 4096, 4097, 8192, 4096 + 4032, 4096 + 4033):
            bits = np.zeros(n, np.uint8)
            bits[rng.choice(n, m, replace=False)] = 1
            w = cases.pack_bits(bits)
            chk = mk.bv(w, n)
            with pkg.BitVector(w, n) as v:
                for what in (3, 4):
                    assert v.serialize(what) == chk.serialize(what), (n, m, what)


def test_wavelet_trees(pkg, oracle, orc):
    mk = _maker(oracle, orc)
    for name, t in texts.text_catalogue(large=True):
        with pkg.WtHuff(t) as wt:
            blob = wt.serialize()
        assert blob == mk.wt_huff(t).serialize(), name
    for name, seq in sequences():
        with pkg.WtInt(seq) as wt:
            assert wt.serialize() == mk.wt_int(seq).serialize(), name
    with pkg.WtHuff(b"") as wt:  # the reference's empty tree serialises uninitialised tables: only the round trip is defined
        with pkg.load_sdsl(wt.serialize(), pkg.KIND_WT_HUFF) as back:
            assert back.size == 0


def test_csa_blob_is_the_reference_blob(pkg, oracle, orc):
    mk = _maker(oracle, orc)
    rng = np.random.default_rng(9)
    for name, t in texts.text_catalogue(zero_free=True, large=True):
        chk = mk.csa(t)
        with pkg.CsaWt(t) as csa:
            blob = csa.serialize()
        assert blob == chk.serialize(), name
        # ... and what the reference loads from it answers like the index it was stored from
        if orc.ref_available() and len(t) > 20:
            loaded = orc.Ref().csa(blob=blob)
            pats = [t[s : s + 6] for s in rng.integers(0, len(t) - 6, 200)]
            flat, off = pkg.csr_patterns(pats)
            assert (loaded.count(flat, off) == chk.count(flat, off)).all(), name


def test_round_trip_through_own_loader(pkg):
    rng = np.random.default_rng(4)
    t = rng.integers(1, 200, 300000, dtype=np.uint8).tobytes()
    pats = [t[s : s + 9] for s in rng.integers(0, len(t) - 9, 500)]
    flat, off = pkg.csr_patterns(pats)
    with pkg.CsaWt(t, sa_dens=8, isa_dens=16) as a:
        blob = a.serialize()
        with pkg.load_sdsl(blob, pkg.KIND_CSA_WT, param=8) as b:
            assert (a.count(flat, off) == b.count(flat, off)).all()
            x, y = a.locate(flat, off), b.locate(flat, off)
            assert (x[0] == y[0]).all() and (x[1] == y[1]).all()
            assert b.serialize() == blob
    with pkg.CsaWt(t, flags=pkg.F_RRR_BV) as a:  # csa_wt<wt_huff<rrr_vector<63>>>
        with pkg.load_sdsl(a.serialize(), pkg.KIND_CSA_WT, flags=pkg.F_RRR_BV) as b:
            assert (a.count(flat, off) == b.count(flat, off)).all()


def _deep_text(rng):
    """symbol k occurs 2^k times: a Huffman tree of depth 17 (one new depth per symbol), shuffled"""
    t = np.concatenate([np.full(1 << k, 65 + k, np.uint8) for k in range(18)])
    rng.shuffle(t)
    return t.tobytes()


def test_device_and_host_builders_agree(pkg, oracle, orc, monkeypatch):
    """wt_build.cu (one stable radix pass per tree depth, on the device) against the host fill (SDSLGPU_HOST_WT=1):
    identical serialised trees, for byte and integer alphabets, plain and rrr-compressed bit vectors"""
    rng = np.random.default_rng(21)
    deep = _deep_text(rng)
    assert pkg.WtHuff(deep).serialize() == _maker(oracle, orc).wt_huff(deep).serialize()
    cases_t = [("deep", deep)] + [(n, t) for n, t in texts.text_catalogue(large=False)]
    for name, t in cases_t:
        for flags in (pkg.F_DEFAULT, pkg.F_RRR_BV):
            monkeypatch.delenv("SDSLGPU_HOST_WT", raising=False)
            with pkg.WtHuff(t, flags=flags) as a:
                x = a.serialize()
            monkeypatch.setenv("SDSLGPU_HOST_WT", "1")
            with pkg.WtHuff(t, flags=flags) as b:
                y = b.serialize()
            assert x == y, (name, flags)
    for name, seq in sequences():
        monkeypatch.delenv("SDSLGPU_HOST_WT", raising=False)
        with pkg.WtInt(seq) as a:
            x, sa = a.serialize(), a.sigma
        monkeypatch.setenv("SDSLGPU_HOST_WT", "1")
        with pkg.WtInt(seq) as b:
            assert x == b.serialize() and sa == b.sigma, name
    monkeypatch.delenv("SDSLGPU_HOST_WT", raising=False)


def test_reference_count_benchmark_index_both_ways(pkg, orc):
    """FM_HUFF of the reference's own count benchmark (benchmark/indexing_count/index.config:8):
    csa_wt<wt_huff<bit_vector, rank_support_v5<>, select_support_scan<>, select_support_scan<0>>, 1<<20, 1<<20>.
    Built here -> the reference's bytes, loadable and countable by the reference; built by the reference -> ingested
    here (SDSLGPU_F_V5_SCAN) with the same counts"""
    if not orc.ref_available():
        pytest.skip("needs oracle/_ref/libsdslref.so (the unmodified reference)")
    R = orc.Ref()
    rng = np.random.default_rng(31)
    dens = 1 << 20
    for name, t in texts.text_catalogue(zero_free=True, large=True):
        ref_blob, ref_count = R.fm_huff(text=t)
        pats = [t[s : s + int(rng.integers(1, 10))] for s in rng.integers(0, max(1, len(t) - 10), 300)] + [b"", b"\x01\x02zz"]
        flat, off = pkg.csr_patterns(pats)
        want = ref_count(flat, off)
        with pkg.CsaWt(t, sa_dens=dens, isa_dens=dens) as csa:
            blob = csa.serialize(1)
            assert blob == ref_blob, name
            assert (csa.count(flat, off) == want).all(), name
        _, loaded_count = R.fm_huff(blob=blob)
        assert (loaded_count(flat, off) == want).all(), (name, "reference on our blob")
        with pkg.load_sdsl(ref_blob, pkg.KIND_CSA_WT, flags=pkg.F_V5_SCAN, param=dens) as back:
            assert (back.count(flat, off) == want).all(), (name, "ingest")
            if len(t) < 10_000:  # one SA sample per 2^20 entries: every occurrence walks back to SA[0]
                a = back.locate(flat[: int(off[10])], off[:11])
                with pkg.CsaWt(t) as plain:
                    b = plain.locate(flat[: int(off[10])], off[:11])
                assert (a[0] == b[0]).all() and (a[1] == b[1]).all(), (name, "locate after ingest")
        with pkg.WtHuff(t) as wt:
            assert wt.serialize(1) == R.wt_huff_v5_blob(t), (name, "wt_huff over rank_support_v5")
