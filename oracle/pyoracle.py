"""oracle/pyoracle.py — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes bindings for the two checkers:
  * ``Oracle``  -> oracle/liboracle.so      (plain-C restatement, oracle*.c)
  * ``Ref``     -> oracle/_ref/libsdslref.so (the unmodified reference headers, ref_driver.cpp)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libsdslref.so")

u64p = C.POINTER(C.c_uint64)
u8p = C.POINTER(C.c_uint8)


def build(verbose=False):
    """Compile liboracle.so (always) and _ref/libsdslref.so (only where /root/reference exists)."""
    r = subprocess.run(["make", "-C", HERE, "all"], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:], r.stderr[-4000:])
    if r.returncode != 0:
        raise RuntimeError("oracle build failed")


def _p64(a):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(u64p)


def _p8(a):
    assert a.dtype == np.uint8 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(u8p)


def _u64(a):
    return np.ascontiguousarray(a, dtype=np.uint64)


def _u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


def _padded_words(words, nbits):
    """copy of the bit vector words with one zero pad word (SDSL allocates +1 word, a2)"""
    nw = (nbits + 63) >> 6
    w = np.zeros(nw + 1, dtype=np.uint64)
    w[:nw] = np.asarray(words, dtype=np.uint64)[:nw]
    return w


def _blob(fn, *args):
    n = fn(*args, None, 0)
    buf = np.zeros(max(int(n), 1), dtype=np.uint8)
    fn(*args, _p8(buf), n)
    return buf[: int(n)].tobytes()


# ------------------------------------------------------------------------------------------------
class Oracle:
    """plain-C restatement"""

    def __init__(self):
        if not os.path.exists(ORACLE_SO):
            build()
        L = self.L = C.CDLL(ORACLE_SO)
        L.orc_cnt.restype = C.c_uint32
        L.orc_cnt.argtypes = [C.c_uint64]
        L.orc_sel.restype = C.c_uint32
        L.orc_sel.argtypes = [C.c_uint64, C.c_uint32]
        L.orc_hi.restype = C.c_uint32
        L.orc_hi.argtypes = [C.c_uint64]
        L.orc_lo.restype = C.c_uint32
        L.orc_lo.argtypes = [C.c_uint64]
        L.orc_rank_v_table_words.restype = C.c_uint64
        L.orc_rank_v_table_words.argtypes = [C.c_uint64]
        L.orc_rank_v_build.restype = None
        L.orc_rank_v_build.argtypes = [u64p, C.c_uint64, C.c_int, u64p]
        L.orc_rank_v_batch.restype = None
        L.orc_rank_v_batch.argtypes = [u64p, u64p, C.c_int, u64p, C.c_uint64, u64p]
        L.orc_rank_v_serialize.restype = C.c_uint64
        L.orc_rank_v_serialize.argtypes = [u64p, C.c_uint64, u8p, C.c_uint64]
        L.orc_rank_v5_table_words.restype = C.c_uint64
        L.orc_rank_v5_table_words.argtypes = [C.c_uint64]
        L.orc_rank_v5_build.restype = None
        L.orc_rank_v5_build.argtypes = [u64p, C.c_uint64, C.c_int, u64p]
        L.orc_rank_v5.restype = C.c_uint64
        L.orc_rank_v5.argtypes = [u64p, u64p, C.c_int, C.c_uint64]
        L.orc_rank_v5_serialize.restype = C.c_uint64
        L.orc_rank_v5_serialize.argtypes = [u64p, C.c_uint64, u8p, C.c_uint64]
        L.orc_select_mcl_build.restype = C.c_void_p
        L.orc_select_mcl_build.argtypes = [u64p, C.c_uint64, C.c_int]
        L.orc_select_mcl_free.restype = None
        L.orc_select_mcl_free.argtypes = [C.c_void_p]
        L.orc_select_mcl_batch.restype = None
        L.orc_select_mcl_batch.argtypes = [C.c_void_p, u64p, u64p, C.c_uint64, u64p]
        L.orc_select_mcl_serialize.restype = C.c_uint64
        L.orc_select_mcl_serialize.argtypes = [C.c_void_p, u8p, C.c_uint64]
        L.orc_bv_serialize.restype = C.c_uint64
        L.orc_bv_serialize.argtypes = [u64p, C.c_uint64, u8p, C.c_uint64]
        L.orc_wt_huff_build.restype = C.c_void_p
        L.orc_wt_huff_build.argtypes = [u8p, C.c_uint64]
        L.orc_wt_huff_free.argtypes = [C.c_void_p]
        for f in (L.orc_wt_huff_rank_batch, L.orc_wt_huff_select_batch):
            f.restype = None
            f.argtypes = [C.c_void_p, u64p, u8p, C.c_uint64, u64p]
        L.orc_wt_huff_access_batch.restype = None
        L.orc_wt_huff_access_batch.argtypes = [C.c_void_p, u64p, C.c_uint64, u64p, u64p]
        L.orc_wt_huff_serialize.restype = C.c_uint64
        L.orc_wt_huff_serialize.argtypes = [C.c_void_p, u8p, C.c_uint64]

        L.orc_csa_build.restype = C.c_void_p
        L.orc_csa_build.argtypes = [u8p, C.c_uint64]
        L.orc_csa_free.argtypes = [C.c_void_p]
        L.orc_csa_count_batch.restype = None
        L.orc_csa_count_batch.argtypes = [C.c_void_p, u8p, u64p, C.c_uint64, u64p, u64p]
        L.orc_csa_locate_batch.restype = None
        L.orc_csa_locate_batch.argtypes = [C.c_void_p, u8p, u64p, C.c_uint64, u64p, u64p]
        L.orc_csa_sa_batch.restype = None
        L.orc_csa_sa_batch.argtypes = [C.c_void_p, u64p, C.c_uint64, u64p]
        L.orc_csa_serialize.restype = C.c_uint64
        L.orc_csa_serialize.argtypes = [C.c_void_p, u8p, C.c_uint64]
        L.orc_csa_extract_batch.restype = None
        L.orc_csa_extract_batch.argtypes = [C.c_void_p, u64p, u64p, C.c_uint64, u64p, u8p]

        for kind in ("rrr", "sd"):
            getattr(L, f"orc_{kind}_build").restype = C.c_void_p
            getattr(L, f"orc_{kind}_build").argtypes = [u64p, C.c_uint64]
            getattr(L, f"orc_{kind}_free").argtypes = [C.c_void_p]
            for op in ("rank", "select"):
                f = getattr(L, f"orc_{kind}_{op}_batch")
                f.restype = None
                f.argtypes = [C.c_void_p, C.c_int, u64p, C.c_uint64, u64p]
            f = getattr(L, f"orc_{kind}_access_batch")
            f.restype = None
            f.argtypes = [C.c_void_p, u64p, C.c_uint64, u64p]
            f = getattr(L, f"orc_{kind}_serialize")
            f.restype = C.c_uint64
            f.argtypes = [C.c_void_p, u8p, C.c_uint64]

        L.orc_wt_int_build.restype = C.c_void_p
        L.orc_wt_int_build.argtypes = [u64p, C.c_uint64]
        L.orc_wt_int_free.argtypes = [C.c_void_p]
        for f in (L.orc_wt_int_rank_batch, L.orc_wt_int_select_batch):
            f.restype = None
            f.argtypes = [C.c_void_p, u64p, u64p, C.c_uint64, u64p]
        L.orc_wt_int_access_batch.restype = None
        L.orc_wt_int_access_batch.argtypes = [C.c_void_p, u64p, C.c_uint64, u64p, u64p]
        L.orc_wt_int_serialize.restype = C.c_uint64
        L.orc_wt_int_serialize.argtypes = [C.c_void_p, u8p, C.c_uint64]

    def csa(self, text):
        return OracleCsa(self, text)

    def wt_int(self, seq):
        return OracleWtInt(self, seq)

    def rrr(self, words, nbits):
        return OracleCompressed(self, "rrr", words, nbits)

    def sd(self, words, nbits):
        return OracleCompressed(self, "sd", words, nbits)

    # -- plain bit vector ------------------------------------------------------------------
    def bv(self, words, nbits):
        return OracleBV(self, words, nbits)

    def wt_huff(self, text):
        return OracleWtHuff(self, text)


class OracleBV:
    def __init__(self, o, words, nbits):
        self.o, self.L = o, o.L
        self.nbits = int(nbits)
        self.w = _padded_words(words, nbits)
        self.tables = {}
        self.sel = {}

    def rank_table(self, b):
        if b not in self.tables:
            t = np.zeros(int(self.L.orc_rank_v_table_words(self.nbits)), dtype=np.uint64)
            self.L.orc_rank_v_build(_p64(self.w), self.nbits, b, _p64(t))
            self.tables[b] = t
        return self.tables[b]

    def rank(self, idx, b=1):
        idx = _u64(idx)
        out = np.zeros(len(idx), dtype=np.uint64)
        self.L.orc_rank_v_batch(_p64(self.w), _p64(self.rank_table(b)), b, _p64(idx), len(idx), _p64(out))
        return out

    def _sel(self, b):
        if b not in self.sel:
            self.sel[b] = self.L.orc_select_mcl_build(_p64(self.w), self.nbits, b)
        return self.sel[b]

    def select(self, i, b=1):
        i = _u64(i)
        out = np.zeros(len(i), dtype=np.uint64)
        self.L.orc_select_mcl_batch(self._sel(b), _p64(self.w), _p64(i), len(i), _p64(out))
        return out

    def rank_v5_table(self, b):
        t = np.zeros(int(self.L.orc_rank_v5_table_words(self.nbits)), dtype=np.uint64)
        self.L.orc_rank_v5_build(_p64(self.w), self.nbits, b, _p64(t))
        return t

    def rank_v5(self, idx, b=1):
        """rank_support_v5<b>::rank, one query at a time (small inputs only)"""
        t = self.rank_v5_table(b)
        return np.array([self.L.orc_rank_v5(_p64(self.w), _p64(t), b, int(i)) for i in idx], dtype=np.uint64)

    def serialize(self, what):
        """what: 0 bit_vector, 1 rank_support_v<1>, 2 rank_support_v<0>, 3 select_support_mcl<1>, 4 <0>,
        5 rank_support_v5<1>, 6 rank_support_v5<0>"""
        if what == 0:
            return _blob(self.L.orc_bv_serialize, _p64(self.w), self.nbits)
        if what in (5, 6):
            return _blob(self.L.orc_rank_v5_serialize, _p64(self.rank_v5_table(1 if what == 5 else 0)), self.nbits)
        if what in (1, 2):
            return _blob(self.L.orc_rank_v_serialize, _p64(self.rank_table(1 if what == 1 else 0)), self.nbits)
        return _blob(self.L.orc_select_mcl_serialize, self._sel(1 if what == 3 else 0))

    def __del__(self):
        for h in self.sel.values():
            self.L.orc_select_mcl_free(h)
        self.sel = {}


def _text(text):
    return _u8(np.frombuffer(text, dtype=np.uint8) if isinstance(text, (bytes, bytearray)) else text)


class OracleWtHuff:
    def __init__(self, o, text):
        self.L = o.L
        t = _text(text)
        self.size = len(t)
        self.h = self.L.orc_wt_huff_build(_p8(t if len(t) else np.zeros(1, np.uint8)), len(t))

    def rank(self, i, c):
        i, c = _u64(i), _u8(c)
        out = np.zeros(len(i), dtype=np.uint64)
        self.L.orc_wt_huff_rank_batch(self.h, _p64(i), _p8(c), len(i), _p64(out))
        return out

    def select(self, i, c):
        i, c = _u64(i), _u8(c)
        out = np.zeros(len(i), dtype=np.uint64)
        self.L.orc_wt_huff_select_batch(self.h, _p64(i), _p8(c), len(i), _p64(out))
        return out

    def inverse_select(self, i):
        i = _u64(i)
        sym = np.zeros(len(i), dtype=np.uint64)
        rnk = np.zeros(len(i), dtype=np.uint64)
        self.L.orc_wt_huff_access_batch(self.h, _p64(i), len(i), _p64(sym), _p64(rnk))
        return rnk, sym

    def access(self, i):
        return self.inverse_select(i)[1]

    def serialize(self):
        return _blob(self.L.orc_wt_huff_serialize, self.h)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_wt_huff_free(self.h)
            self.h = None


class OracleWtInt:
    def __init__(self, o, seq):
        self.L = o.L
        s = _u64(seq)
        self.size = len(s)
        self.h = self.L.orc_wt_int_build(_p64(s if len(s) else np.zeros(1, np.uint64)), len(s))

    def rank(self, i, c):
        i, c = _u64(i), _u64(c)
        out = np.zeros(len(i), dtype=np.uint64)
        self.L.orc_wt_int_rank_batch(self.h, _p64(i), _p64(c), len(i), _p64(out))
        return out

    def select(self, i, c):
        i, c = _u64(i), _u64(c)
        out = np.zeros(len(i), dtype=np.uint64)
        self.L.orc_wt_int_select_batch(self.h, _p64(i), _p64(c), len(i), _p64(out))
        return out

    def inverse_select(self, i):
        i = _u64(i)
        sym = np.zeros(len(i), dtype=np.uint64)
        rnk = np.zeros(len(i), dtype=np.uint64)
        self.L.orc_wt_int_access_batch(self.h, _p64(i), len(i), _p64(sym), _p64(rnk))
        return rnk, sym

    def access(self, i):
        return self.inverse_select(i)[1]

    def serialize(self):
        return _blob(self.L.orc_wt_int_serialize, self.h)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_wt_int_free(self.h)
            self.h = None


class OracleCompressed:
    """rrr_vector<63> / sd_vector<> restatement"""

    def __init__(self, o, kind, words, nbits):
        self.L, self.kind = o.L, kind
        self.nbits = int(nbits)
        self.w = _padded_words(words, nbits)
        self.h = getattr(self.L, f"orc_{kind}_build")(_p64(self.w), self.nbits)

    def _q(self, op, x, b):
        x = _u64(x)
        out = np.zeros(len(x), dtype=np.uint64)
        getattr(self.L, f"orc_{self.kind}_{op}_batch")(self.h, b, _p64(x), len(x), _p64(out))
        return out

    def rank(self, idx, b=1):
        return self._q("rank", idx, b)

    def select(self, i, b=1):
        return self._q("select", i, b)

    def access(self, idx):
        idx = _u64(idx)
        out = np.zeros(len(idx), dtype=np.uint64)
        getattr(self.L, f"orc_{self.kind}_access_batch")(self.h, _p64(idx), len(idx), _p64(out))
        return out

    def serialize(self):
        return _blob(getattr(self.L, f"orc_{self.kind}_serialize"), self.h)

    def __del__(self):
        if getattr(self, "h", None):
            getattr(self.L, f"orc_{self.kind}_free")(self.h)
            self.h = None


class OracleCsa:
    def __init__(self, o, text):
        self.L = o.L
        t = _text(text)
        assert not (t == 0).any(), "csa texts must be zero-free (construct.hpp:34-46)"
        self.size = len(t) + 1
        self.h = self.L.orc_csa_build(_p8(t if len(t) else np.zeros(1, np.uint8)), len(t))

    def count(self, flat, off, want_l=False):
        n = len(off) - 1
        cnt = np.zeros(n, dtype=np.uint64)
        l = np.zeros(n, dtype=np.uint64) if want_l else None
        self.L.orc_csa_count_batch(self.h, _p8(flat), _p64(off), n, _p64(cnt), _p64(l) if want_l else None)
        return (cnt, l) if want_l else cnt

    def locate(self, flat, off):
        cnt = self.count(flat, off)
        occ_off = np.zeros(len(cnt) + 1, dtype=np.uint64)
        occ_off[1:] = np.cumsum(cnt, dtype=np.uint64)
        occ = np.zeros(max(int(occ_off[-1]), 1), dtype=np.uint64)
        self.L.orc_csa_locate_batch(self.h, _p8(flat), _p64(off), len(cnt), _p64(occ_off), _p64(occ))
        return occ_off, occ[: int(occ_off[-1])]

    def sa(self, i):
        i = _u64(i)
        out = np.zeros(len(i), dtype=np.uint64)
        self.L.orc_csa_sa_batch(self.h, _p64(i), len(i), _p64(out))
        return out

    def serialize(self):
        return _blob(self.L.orc_csa_serialize, self.h)

    def extract(self, begin, end):
        """text[begin[k] .. end[k]] (inclusive) for every k -> (offsets uint64[n+1], bytes)"""
        begin, end = _u64(begin), _u64(end)
        off = np.zeros(len(begin) + 1, dtype=np.uint64)
        off[1:] = np.cumsum(end - begin + np.uint64(1), dtype=np.uint64)
        out = np.zeros(max(int(off[-1]), 1), dtype=np.uint8)
        self.L.orc_csa_extract_batch(self.h, _p64(begin), _p64(end), len(begin), _p64(off), _p8(out))
        return off, out[: int(off[-1])]

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_csa_free(self.h)
            self.h = None


# ------------------------------------------------------------------------------------------------
def ref_available():
    return os.path.exists(REF_SO)


class Ref:
    """the unmodified reference (sdsl-lite 3.0.5), through oracle/ref_driver.cpp"""

    def __init__(self):
        if not os.path.exists(REF_SO):
            raise RuntimeError("oracle/_ref/libsdslref.so missing (run make -C oracle where /root/reference exists)")
        L = self.L = C.CDLL(REF_SO)
        vp = C.c_void_p
        L.ref_bv_create.restype = vp
        L.ref_bv_create.argtypes = [u64p, C.c_uint64, C.c_int]
        L.ref_bv_free.argtypes = [vp]
        for f in (L.ref_bv_rank, L.ref_bv_select):
            f.restype = None
            f.argtypes = [vp, C.c_int, u64p, C.c_uint64, u64p, C.c_int]
        L.ref_bv_serialize.restype = C.c_uint64
        L.ref_bv_serialize.argtypes = [vp, C.c_int, u8p, C.c_uint64]
        for kind in ("rrr", "sd"):
            getattr(L, f"ref_{kind}_create").restype = vp
            getattr(L, f"ref_{kind}_create").argtypes = [u64p, C.c_uint64]
            getattr(L, f"ref_{kind}_free").argtypes = [vp]
            for op in ("rank", "select"):
                f = getattr(L, f"ref_{kind}_{op}")
                f.restype = None
                f.argtypes = [vp, C.c_int, u64p, C.c_uint64, u64p, C.c_int]
            f = getattr(L, f"ref_{kind}_access")
            f.restype = None
            f.argtypes = [vp, u64p, C.c_uint64, u64p, C.c_int]
            f = getattr(L, f"ref_{kind}_serialize")
            f.restype = C.c_uint64
            f.argtypes = [vp, u8p, C.c_uint64]
        L.ref_wt_huff_create.restype = vp
        L.ref_wt_huff_create.argtypes = [u8p, C.c_uint64]
        L.ref_wt_huff_free.argtypes = [vp]
        L.ref_wt_huff_size.restype = C.c_uint64
        L.ref_wt_huff_size.argtypes = [vp]
        L.ref_wt_huff_sigma.restype = C.c_uint64
        L.ref_wt_huff_sigma.argtypes = [vp]
        for f in (L.ref_wt_huff_rank, L.ref_wt_huff_select):
            f.restype = None
            f.argtypes = [vp, u64p, u8p, C.c_uint64, u64p, C.c_int]
        L.ref_wt_huff_access.restype = None
        L.ref_wt_huff_access.argtypes = [vp, u64p, C.c_uint64, u64p, u64p, C.c_int]
        L.ref_wt_huff_serialize.restype = C.c_uint64
        L.ref_wt_huff_serialize.argtypes = [vp, u8p, C.c_uint64]
        L.ref_wt_int_create.restype = vp
        L.ref_wt_int_create.argtypes = [u64p, C.c_uint64]
        L.ref_wt_int_free.argtypes = [vp]
        L.ref_wt_int_sigma.restype = C.c_uint64
        L.ref_wt_int_sigma.argtypes = [vp]
        L.ref_wt_int_max_level.restype = C.c_uint64
        L.ref_wt_int_max_level.argtypes = [vp]
        for f in (L.ref_wt_int_rank, L.ref_wt_int_select):
            f.restype = None
            f.argtypes = [vp, u64p, u64p, C.c_uint64, u64p, C.c_int]
        L.ref_wt_int_access.restype = None
        L.ref_wt_int_access.argtypes = [vp, u64p, C.c_uint64, u64p, u64p, C.c_int]
        L.ref_wt_int_serialize.restype = C.c_uint64
        L.ref_wt_int_serialize.argtypes = [vp, u8p, C.c_uint64]
        L.ref_csa_create.restype = vp
        L.ref_csa_create.argtypes = [u8p, C.c_uint64]
        L.ref_csa_load.restype = vp
        L.ref_csa_load.argtypes = [u8p, C.c_uint64]
        L.ref_csa_free.argtypes = [vp]
        L.ref_csa_size.restype = C.c_uint64
        L.ref_csa_size.argtypes = [vp]
        L.ref_csa_serialize.restype = C.c_uint64
        L.ref_csa_serialize.argtypes = [vp, u8p, C.c_uint64]
        L.ref_csa_count.restype = None
        L.ref_csa_count.argtypes = [vp, u8p, u64p, C.c_uint64, u64p, u64p, C.c_int]
        L.ref_csa_locate.restype = None
        L.ref_csa_locate.argtypes = [vp, u8p, u64p, C.c_uint64, u64p, u64p, C.c_int]
        L.ref_csa_sa.restype = None
        L.ref_csa_sa.argtypes = [vp, u64p, C.c_uint64, u64p, C.c_int]
        L.ref_csa_bwt_rank.restype = None
        L.ref_csa_bwt_rank.argtypes = [vp, u64p, u8p, C.c_uint64, u64p, C.c_int]
        L.ref_csa_extract.restype = None
        L.ref_csa_extract.argtypes = [vp, C.c_uint64, C.c_uint64, u8p]
        L.ref_version.restype = C.c_char_p
        if hasattr(L, "ref_wt_huff_rrr_create"):
            L.ref_wt_huff_rrr_create.restype = vp
            L.ref_wt_huff_rrr_create.argtypes = [u8p, C.c_uint64]
            L.ref_wt_huff_rrr_free.argtypes = [vp]
            L.ref_wt_huff_rrr_rank.restype = None
            L.ref_wt_huff_rrr_rank.argtypes = [vp, u64p, u8p, C.c_uint64, u64p, C.c_int]
            L.ref_wt_huff_rrr_serialize.restype = C.c_uint64
            L.ref_wt_huff_rrr_serialize.argtypes = [vp, u8p, C.c_uint64]
            L.ref_csa_rrr_create.restype = vp
            L.ref_csa_rrr_create.argtypes = [u8p, C.c_uint64]
            L.ref_csa_rrr_free.argtypes = [vp]
            L.ref_csa_rrr_serialize.restype = C.c_uint64
            L.ref_csa_rrr_serialize.argtypes = [vp, u8p, C.c_uint64]
            L.ref_csa_rrr_count.restype = None
            L.ref_csa_rrr_count.argtypes = [vp, u8p, u64p, C.c_uint64, u64p, C.c_int]

        if hasattr(L, "ref_fm_huff_create"):
            for name in ("ref_wt_huff_v5_create", "ref_fm_huff_create", "ref_fm_huff_load"):
                getattr(L, name).restype = vp
                getattr(L, name).argtypes = [u8p, C.c_uint64]
            L.ref_wt_huff_v5_free.argtypes = [vp]
            L.ref_fm_huff_free.argtypes = [vp]
            for name in ("ref_wt_huff_v5_serialize", "ref_fm_huff_serialize"):
                getattr(L, name).restype = C.c_uint64
                getattr(L, name).argtypes = [vp, u8p, C.c_uint64]
            L.ref_fm_huff_count.restype = None
            L.ref_fm_huff_count.argtypes = [vp, u8p, u64p, C.c_uint64, u64p, C.c_int]

    def wt_huff_v5_blob(self, text):
        """serialised wt_huff<bit_vector, rank_support_v5<>, select_support_scan<>, select_support_scan<0>> of `text`"""
        t = _text(text)
        h = self.L.ref_wt_huff_v5_create(_p8(t if len(t) else np.zeros(1, np.uint8)), len(t))
        blob = _blob(self.L.ref_wt_huff_v5_serialize, h)
        self.L.ref_wt_huff_v5_free(h)
        return blob

    def fm_huff(self, text=None, blob=None):
        """the reference's count-benchmark index FM_HUFF (benchmark/indexing_count/index.config:8), built from `text`
        or loaded from `blob` -> (serialised bytes, count_fn)"""
        if blob is not None:
            b = np.frombuffer(blob, dtype=np.uint8)
            h = self.L.ref_fm_huff_load(_p8(b), len(b))
        else:
            t = _text(text)
            h = self.L.ref_fm_huff_create(_p8(t), len(t))
        out_blob = _blob(self.L.ref_fm_huff_serialize, h)

        def count(flat, off):
            out = np.zeros(len(off) - 1, dtype=np.uint64)
            self.L.ref_fm_huff_count(h, _p8(flat), _p64(off), len(off) - 1, _p64(out), 1)
            return out

        return out_blob, count

    def bv(self, words, nbits, with_select=True):
        return RefBV(self, words, nbits, with_select)

    def rrr(self, words, nbits):
        return RefCompressed(self, "rrr", words, nbits)

    def sd(self, words, nbits):
        return RefCompressed(self, "sd", words, nbits)

    def wt_huff(self, text):
        return RefWtHuff(self, text)

    def wt_int(self, seq):
        return RefWtInt(self, seq)

    def csa(self, text=None, blob=None):
        return RefCsa(self, text, blob)

    def wt_huff_rrr_blob(self, text):
        """serialised wt_huff<rrr_vector<63>> of `text` and its rank answers for (i, c) -> (blob, rank_fn)"""
        t = _text(text)
        h = self.L.ref_wt_huff_rrr_create(_p8(t if len(t) else np.zeros(1, np.uint8)), len(t))
        blob = _blob(self.L.ref_wt_huff_rrr_serialize, h)

        def rank(i, c):
            i, c = _u64(i), _u8(c)
            out = np.zeros(len(i), dtype=np.uint64)
            self.L.ref_wt_huff_rrr_rank(h, _p64(i), _p8(c), len(i), _p64(out), 1)
            return out

        return blob, rank

    def csa_rrr_blob(self, text):
        """serialised csa_wt<wt_huff<rrr_vector<63>>> of `text` and its count() -> (blob, count_fn)"""
        t = _text(text)
        h = self.L.ref_csa_rrr_create(_p8(t if len(t) else np.zeros(1, np.uint8)), len(t))
        blob = _blob(self.L.ref_csa_rrr_serialize, h)

        def count(flat, off):
            out = np.zeros(len(off) - 1, dtype=np.uint64)
            self.L.ref_csa_rrr_count(h, _p8(flat), _p64(off), len(off) - 1, _p64(out), 1)
            return out

        return blob, count


class RefBV:
    def __init__(self, r, words, nbits, with_select):
        self.L = r.L
        self.nbits = int(nbits)
        w = _padded_words(words, nbits)
        self.h = self.L.ref_bv_create(_p64(w), self.nbits, 1 if with_select else 0)

    def rank(self, idx, b=1, threads=1):
        idx = _u64(idx)
        out = np.zeros(len(idx), dtype=np.uint64)
        self.L.ref_bv_rank(self.h, b, _p64(idx), len(idx), _p64(out), threads)
        return out

    def select(self, i, b=1, threads=1):
        i = _u64(i)
        out = np.zeros(len(i), dtype=np.uint64)
        self.L.ref_bv_select(self.h, b, _p64(i), len(i), _p64(out), threads)
        return out

    def serialize(self, what):
        return _blob(self.L.ref_bv_serialize, self.h, what)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_bv_free(self.h)
            self.h = None


class RefCompressed:
    def __init__(self, r, kind, words, nbits):
        self.L, self.kind = r.L, kind
        self.nbits = int(nbits)
        w = _padded_words(words, nbits)
        self.h = getattr(self.L, f"ref_{kind}_create")(_p64(w), self.nbits)

    def _q(self, op, x, b, threads):
        x = _u64(x)
        out = np.zeros(len(x), dtype=np.uint64)
        getattr(self.L, f"ref_{self.kind}_{op}")(self.h, b, _p64(x), len(x), _p64(out), threads)
        return out

    def rank(self, idx, b=1, threads=1):
        return self._q("rank", idx, b, threads)

    def select(self, i, b=1, threads=1):
        return self._q("select", i, b, threads)

    def access(self, idx, threads=1):
        idx = _u64(idx)
        out = np.zeros(len(idx), dtype=np.uint64)
        getattr(self.L, f"ref_{self.kind}_access")(self.h, _p64(idx), len(idx), _p64(out), threads)
        return out

    def serialize(self):
        return _blob(getattr(self.L, f"ref_{self.kind}_serialize"), self.h)

    def __del__(self):
        if getattr(self, "h", None):
            getattr(self.L, f"ref_{self.kind}_free")(self.h)
            self.h = None


class RefWtHuff:
    def __init__(self, r, text):
        self.L = r.L
        t = _u8(np.frombuffer(text, dtype=np.uint8) if isinstance(text, (bytes, bytearray)) else text)
        self.h = self.L.ref_wt_huff_create(_p8(t), len(t))
        self.size = int(self.L.ref_wt_huff_size(self.h))
        self.sigma = int(self.L.ref_wt_huff_sigma(self.h))

    def rank(self, i, c, threads=1):
        i, c = _u64(i), _u8(c)
        out = np.zeros(len(i), dtype=np.uint64)
        self.L.ref_wt_huff_rank(self.h, _p64(i), _p8(c), len(i), _p64(out), threads)
        return out

    def select(self, i, c, threads=1):
        i, c = _u64(i), _u8(c)
        out = np.zeros(len(i), dtype=np.uint64)
        self.L.ref_wt_huff_select(self.h, _p64(i), _p8(c), len(i), _p64(out), threads)
        return out

    def access(self, i, threads=1):
        i = _u64(i)
        out = np.zeros(len(i), dtype=np.uint64)
        self.L.ref_wt_huff_access(self.h, _p64(i), len(i), _p64(out), None, threads)
        return out

    def inverse_select(self, i, threads=1):
        i = _u64(i)
        sym = np.zeros(len(i), dtype=np.uint64)
        rnk = np.zeros(len(i), dtype=np.uint64)
        self.L.ref_wt_huff_access(self.h, _p64(i), len(i), _p64(sym), _p64(rnk), threads)
        return rnk, sym

    def serialize(self):
        return _blob(self.L.ref_wt_huff_serialize, self.h)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_wt_huff_free(self.h)
            self.h = None


class RefWtInt:
    def __init__(self, r, seq):
        self.L = r.L
        s = _u64(seq)
        self.h = self.L.ref_wt_int_create(_p64(s), len(s))
        self.size = len(s)
        self.sigma = int(self.L.ref_wt_int_sigma(self.h))
        self.max_level = int(self.L.ref_wt_int_max_level(self.h))

    def rank(self, i, c, threads=1):
        i, c = _u64(i), _u64(c)
        out = np.zeros(len(i), dtype=np.uint64)
        self.L.ref_wt_int_rank(self.h, _p64(i), _p64(c), len(i), _p64(out), threads)
        return out

    def select(self, i, c, threads=1):
        i, c = _u64(i), _u64(c)
        out = np.zeros(len(i), dtype=np.uint64)
        self.L.ref_wt_int_select(self.h, _p64(i), _p64(c), len(i), _p64(out), threads)
        return out

    def access(self, i, threads=1):
        i = _u64(i)
        out = np.zeros(len(i), dtype=np.uint64)
        self.L.ref_wt_int_access(self.h, _p64(i), len(i), _p64(out), None, threads)
        return out

    def inverse_select(self, i, threads=1):
        i = _u64(i)
        sym = np.zeros(len(i), dtype=np.uint64)
        rnk = np.zeros(len(i), dtype=np.uint64)
        self.L.ref_wt_int_access(self.h, _p64(i), len(i), _p64(sym), _p64(rnk), threads)
        return rnk, sym

    def serialize(self):
        return _blob(self.L.ref_wt_int_serialize, self.h)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_wt_int_free(self.h)
            self.h = None


def csr_patterns(pats):
    """list of bytes -> (uint8 concatenation, uint64 offsets[n+1])"""
    off = np.zeros(len(pats) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(p) for p in pats], dtype=np.uint64)
    flat = np.frombuffer(b"".join(pats), dtype=np.uint8).copy() if len(pats) else np.zeros(0, np.uint8)
    if len(flat) == 0:
        flat = np.zeros(1, np.uint8)
    return flat, off


class RefCsa:
    def __init__(self, r, text=None, blob=None):
        self.L = r.L
        if blob is not None:
            b = np.frombuffer(blob, dtype=np.uint8).copy()
            self.h = self.L.ref_csa_load(_p8(b), len(b))
        else:
            t = _u8(np.frombuffer(text, dtype=np.uint8) if isinstance(text, (bytes, bytearray)) else text)
            if len(t) == 0:
                t = np.zeros(1, np.uint8)
                self.h = self.L.ref_csa_create(_p8(t), 0)
            else:
                self.h = self.L.ref_csa_create(_p8(t), len(t))
        if not self.h:
            raise ValueError("reference refused the text (contains a zero byte?)")
        self.size = int(self.L.ref_csa_size(self.h))

    def count(self, flat, off, threads=1, want_l=False):
        n = len(off) - 1
        cnt = np.zeros(n, dtype=np.uint64)
        l = np.zeros(n, dtype=np.uint64) if want_l else None
        self.L.ref_csa_count(self.h, _p8(flat), _p64(off), n, _p64(cnt), _p64(l) if want_l else None, threads)
        return (cnt, l) if want_l else cnt

    def locate(self, flat, off, threads=1):
        n = len(off) - 1
        occ_off = np.zeros(n + 1, dtype=np.uint64)
        self.L.ref_csa_locate(self.h, _p8(flat), _p64(off), n, _p64(occ_off), None, threads)
        occ = np.zeros(max(int(occ_off[-1]), 1), dtype=np.uint64)
        self.L.ref_csa_locate(self.h, _p8(flat), _p64(off), n, _p64(occ_off), _p64(occ), threads)
        return occ_off, occ[: int(occ_off[-1])]

    def sa(self, i, threads=1):
        i = _u64(i)
        out = np.zeros(len(i), dtype=np.uint64)
        self.L.ref_csa_sa(self.h, _p64(i), len(i), _p64(out), threads)
        return out

    def bwt_rank(self, i, c, threads=1):
        i, c = _u64(i), _u8(c)
        out = np.zeros(len(i), dtype=np.uint64)
        self.L.ref_csa_bwt_rank(self.h, _p64(i), _p8(c), len(i), _p64(out), threads)
        return out

    def extract(self, lo, hi):
        out = np.zeros(hi - lo + 1, dtype=np.uint8)
        self.L.ref_csa_extract(self.h, lo, hi, _p8(out))
        return out.tobytes()

    def serialize(self):
        return _blob(self.L.ref_csa_serialize, self.h)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_csa_free(self.h)
            self.h = None
