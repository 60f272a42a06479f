"""sdsl-lite_b200 — Python host binding of the C-ABI product library (include/sdslgpu.h).

This file is plumbing for tests/ and bench.py: a ctypes wrapper around ``libsdslgpu.so`` whose
classes mirror the reference's interface for the hot path (``rank(i)``, ``select(i)``,
``wt.rank(i, c)``, ``count``, ``locate`` — SURVEY.md §8(b)), in batch form.  The product itself is
the CUDA library; there is NO CPU fallback: if the shared object is missing or no CUDA device is
present every constructor raises.

Arguments may be numpy arrays (host buffers: the library stages them through the chunked PCIe
pipeline) or torch CUDA tensors (device buffers: the kernel is launched asynchronously on the given
stream / torch's current stream).

The directory name contains a hyphen, so import it with ``__graft_entry__.load_package()`` which
registers it as module ``sdsl_lite_b200``.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SDSLGPU_LIB") or os.path.join(HERE, "libsdslgpu.so")  # SDSLGPU_LIB: a variant build (tools/variants.sh)

OK, EINVAL, ENOMEM, ECUDA, ENOTSUP = 0, -1, -2, -3, -4
NPOS = np.uint64(0xFFFFFFFFFFFFFFFF)
F_DEFAULT, F_SDSL_LAYOUT, F_NO_SELECT, F_RRR_BV, F_COMPACT, F_V5_SCAN = 0, 1, 2, 4, 8, 16
KIND_BV, KIND_RRR63, KIND_SD, KIND_WT_HUFF, KIND_WT_INT, KIND_CSA_WT = 1, 2, 3, 4, 5, 6
PAT_0, PAT_1, PAT_10, PAT_01, PAT_00, PAT_11 = 0, 1, 2, 3, 4, 5  # <t_b, t_pat_len> of rank_support_v / select_support_mcl
ORDER_AUTO, ORDER_DIRECT, ORDER_BINNED = 0, 1, 2  # sdslgpu_set_batch_order

u64p = C.POINTER(C.c_uint64)
vp = C.c_void_p

# every symbol include/sdslgpu.h declares: (name, restype, argtypes)
_SIGNATURES = [
    ("sdslgpu_version", C.c_char_p, []),
    ("sdslgpu_last_error", C.c_char_p, []),
    ("sdslgpu_device_count", C.c_int, [C.POINTER(C.c_int)]),
    ("sdslgpu_bv_create", C.c_int, [vp, C.c_uint64, C.c_int, C.c_uint32, C.POINTER(vp)]),
    ("sdslgpu_free", C.c_int, [vp]),
    ("sdslgpu_kind", C.c_int, [vp, C.POINTER(C.c_int)]),
    ("sdslgpu_size", C.c_int, [vp, u64p]),
    ("sdslgpu_arg_count", C.c_int, [vp, C.c_int, u64p]),
    ("sdslgpu_device_bytes", C.c_int, [vp, u64p]),
    ("sdslgpu_rank", C.c_int, [vp, C.c_int, vp, C.c_uint64, vp, vp]),
    ("sdslgpu_select", C.c_int, [vp, C.c_int, vp, C.c_uint64, vp, vp]),
    ("sdslgpu_access", C.c_int, [vp, vp, C.c_uint64, vp, vp]),
    ("sdslgpu_set_batch_order", C.c_int, [vp, C.c_int]),
    ("sdslgpu_auto_is_binned", C.c_int, [C.c_uint64, C.c_uint64, C.c_int]),
    ("sdslgpu_rank_iv", C.c_int, [vp, C.c_int, vp, C.c_uint32, C.c_uint64, vp, C.c_uint32, vp]),
    ("sdslgpu_select_iv", C.c_int, [vp, C.c_int, vp, C.c_uint32, C.c_uint64, vp, C.c_uint32, vp]),
    ("sdslgpu_bv_serialize", C.c_int, [vp, C.c_int, vp, C.c_uint64, u64p]),
    ("sdslgpu_wt_huff_create", C.c_int, [vp, C.c_uint64, C.c_int, C.c_uint32, C.POINTER(vp)]),
    ("sdslgpu_wt_sigma", C.c_int, [vp, u64p]),
    ("sdslgpu_wt_rank", C.c_int, [vp, vp, vp, C.c_uint64, vp, vp]),
    ("sdslgpu_wt_select", C.c_int, [vp, vp, vp, C.c_uint64, vp, vp]),
    ("sdslgpu_wt_access", C.c_int, [vp, vp, C.c_uint64, vp, vp, vp]),
]

_lib = None


class SdslGpuError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"sdslgpu status {status}: {msg}")
        self.status = status


def build(verbose=False):
    """Compile libsdslgpu.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", HERE, "-j8", "all"], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-6000:], r.stderr[-6000:])
    if r.returncode != 0:
        raise RuntimeError("building libsdslgpu.so failed")


def lib():
    """Load the product library; fails loudly when it is missing (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run __graft_entry__.build() (there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, res, args in _SIGNATURES:
            f = getattr(L, name)  # AttributeError here = the .so does not export a declared symbol
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


def declared_symbols():
    return [s[0] for s in _SIGNATURES]


def _check(st):
    if st != OK:
        raise SdslGpuError(st, lib().sdslgpu_last_error().decode())


def _is_torch(x):
    return type(x).__module__.startswith("torch")


def _in_ptr(x, dtype=np.uint64):
    """-> (pointer, n, keepalive, on_device)"""
    if _is_torch(x):
        import torch

        want = {np.uint64: (torch.int64, torch.uint64), np.uint8: (torch.uint8,)}[dtype]
        assert x.dtype in want and x.is_contiguous(), "torch inputs must be contiguous int64/uint64 (or uint8)"
        return x.data_ptr(), x.numel(), x, x.is_cuda
    a = np.ascontiguousarray(x, dtype=dtype)
    return a.ctypes.data, a.size, a, False


def _out_like(x, n, out=None):
    """allocate (or validate) the uint64 output next to the input: torch -> torch, numpy -> numpy"""
    if out is not None:
        if _is_torch(out):
            return out.data_ptr(), out, out
        assert out.dtype == np.uint64 and out.flags["C_CONTIGUOUS"] and out.size >= n
        return out.ctypes.data, out, out
    if _is_torch(x):
        import torch

        o = torch.empty(n, dtype=torch.int64, device=x.device)
        return o.data_ptr(), o, o
    o = np.empty(n, dtype=np.uint64)
    return o.ctypes.data, o, o


def _stream_ptr(stream, x):
    if stream is not None:
        return int(getattr(stream, "cuda_stream", stream))
    if _is_torch(x) and x.is_cuda:
        import torch

        return int(torch.cuda.current_stream(x.device).cuda_stream)
    return 0


class _Handle:
    def __init__(self):
        self._h = vp()

    def close(self):
        if self._h:
            lib().sdslgpu_free(self._h)
            self._h = vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def size(self):
        v = C.c_uint64()
        _check(lib().sdslgpu_size(self._h, C.byref(v)))
        return v.value

    def __len__(self):
        return self.size

    @property
    def device_bytes(self):
        v = C.c_uint64()
        _check(lib().sdslgpu_device_bytes(self._h, C.byref(v)))
        return v.value

    # ---- the batched bit-vector concept (rank_support / select_support) ----------
    def arg_count(self, b=1):
        v = C.c_uint64()
        _check(lib().sdslgpu_arg_count(self._h, b, C.byref(v)))
        return v.value

    def set_batch_order(self, order):
        """ORDER_AUTO / ORDER_DIRECT / ORDER_BINNED: how large rank / select batches on bit-vector handles (plain, rrr,
        sd) are executed; never changes a result"""
        _check(lib().sdslgpu_set_batch_order(self._h, int(order)))

    def serialize(self, what=0):
        """the reference's serialize() / store_to_file bytes (sdslgpu_serialize): bit vectors what = 0 bit_vector,
        1 / 2 rank_support_v<1> / <0> (need F_SDSL_LAYOUT), 3 / 4 select_support_mcl<1> / <0>; sd_vector what = 1 the
        complete blob; rrr_vector, wavelet trees and csa_wt what = 0 the complete blob"""
        n = C.c_uint64()
        _check(lib().sdslgpu_serialize(self._h, what, None, 0, C.byref(n)))
        buf = np.empty(max(n.value, 1), dtype=np.uint8)
        _check(lib().sdslgpu_serialize(self._h, what, buf.ctypes.data, n.value, C.byref(n)))
        return buf[: n.value].tobytes()

    def rank(self, idx, b=1, out=None, stream=None):
        p, n, keep, _ = _in_ptr(idx)
        po, o, _k = _out_like(idx, n, out)
        _check(lib().sdslgpu_rank(self._h, b, p, n, po, _stream_ptr(stream, idx)))
        return o

    def select(self, i, b=1, out=None, stream=None):
        p, n, keep, _ = _in_ptr(i)
        po, o, _k = _out_like(i, n, out)
        _check(lib().sdslgpu_select(self._h, b, p, n, po, _stream_ptr(stream, i)))
        return o

    def access(self, idx, out=None, stream=None):
        p, n, keep, _ = _in_ptr(idx)
        po, o, _k = _out_like(idx, n, out)
        _check(lib().sdslgpu_access(self._h, p, n, po, _stream_ptr(stream, idx)))
        return o

    # ---- the same with queries / results as int_vector<w> fields (sdslgpu_rank_iv / _select_iv) ----
    def _iv(self, fn, words, width, n, b, out_width, out, stream):
        p, nw, keep, _ = _in_ptr(words)
        assert nw >= iv_words(n, width), "packed query array too short"
        po, o, _k = _out_like(words, iv_words(n, out_width), out)
        _check(fn(self._h, b, p, int(width), int(n), po, int(out_width), _stream_ptr(stream, words)))
        return o

    def rank_iv(self, words, width, n, b=1, out_width=64, out=None, stream=None):
        return self._iv(lib().sdslgpu_rank_iv, words, width, n, b, out_width, out, stream)

    def select_iv(self, words, width, n, b=1, out_width=64, out=None, stream=None):
        return self._iv(lib().sdslgpu_select_iv, words, width, n, b, out_width, out, stream)


def iv_words(n, width):
    return (int(n) * int(width) + 63) >> 6


def iv_pack(values, width):
    """u64 values -> the word array of an int_vector<width> (field k at bits [k*width, (k+1)*width), LSB first).
    Host-side helper for tests / bench (numpy): 64 fields fill exactly `width` words, so the work is 64 vectorised passes."""
    v = np.ascontiguousarray(values, dtype=np.uint64)
    n = len(v)
    mask = np.uint64((1 << width) - 1) if width < 64 else np.uint64(2**64 - 1)
    groups = (n + 63) // 64
    pad = np.zeros(groups * 64, dtype=np.uint64)
    pad[:n] = v & mask
    pad = pad.reshape(groups, 64)
    out = np.zeros((groups, width + 1), dtype=np.uint64)
    for j in range(64):
        pos = j * width
        w, off = pos >> 6, pos & 63
        out[:, w] |= pad[:, j] << np.uint64(off)
        if off + width > 64:
            out[:, w + 1] |= pad[:, j] >> np.uint64(64 - off)
    return out[:, :width].reshape(-1)[: iv_words(n, width)].copy()


def iv_unpack(words, width, n):
    w = np.ascontiguousarray(words, dtype=np.uint64)
    groups = (n + 63) // 64
    need = groups * width
    buf = np.zeros(need + 1, dtype=np.uint64)
    buf[: min(len(w), need)] = w[:need]
    g = buf[:need].reshape(groups, width)
    gx = np.concatenate([g, np.zeros((groups, 1), np.uint64)], axis=1)
    out = np.zeros((groups, 64), dtype=np.uint64)
    mask = np.uint64((1 << width) - 1) if width < 64 else np.uint64(2**64 - 1)
    for j in range(64):
        pos = j * width
        wi, off = pos >> 6, pos & 63
        x = gx[:, wi] >> np.uint64(off)
        if off + width > 64:
            x = x | (gx[:, wi + 1] << np.uint64(64 - off))
        out[:, j] = x & mask
    return out.reshape(-1)[:n].copy()


class PackedBatch:
    """bench.py's e2e leg over the int_vector<w> wire format: the rank and select queries of one step as packed,
    pinned host arrays (w = bits needed for a position / an index), results into packed pinned host arrays"""

    def __init__(self, bv, idx, sel, nbits):
        import torch

        self.bv, self.n = bv, len(idx)
        self.w = max(int(nbits).bit_length(), 1)  # positions 0..nbits and ranks 0..nbits fit
        self.h_idx = torch.from_numpy(iv_pack(idx, self.w).view(np.int64)).pin_memory()
        self.h_sel = torch.from_numpy(iv_pack(sel, self.w).view(np.int64)).pin_memory()
        self.h_out_r = torch.empty(iv_words(self.n, self.w), dtype=torch.int64).pin_memory()
        self.h_out_s = torch.empty(iv_words(self.n, self.w), dtype=torch.int64).pin_memory()
        self.h2d_bytes = 2 * iv_words(self.n, self.w) * 8
        self.d2h_bytes = 2 * iv_words(self.n, self.w) * 8

    def _np(self, t):
        return t.numpy().view(np.uint64)

    def step(self):
        self.bv.rank_iv(self._np(self.h_idx), self.w, self.n, 1, self.w, out=self._np(self.h_out_r))
        self.bv.select_iv(self._np(self.h_sel), self.w, self.n, 1, self.w, out=self._np(self.h_out_s))

    def check(self, bv, idx_prefix, sel_prefix):
        """values of the packed path == values of the u64 path on a prefix"""
        k = len(idx_prefix)
        mask = np.uint64((1 << self.w) - 1)
        r = iv_unpack(self._np(self.h_out_r), self.w, k)
        s_ = iv_unpack(self._np(self.h_out_s), self.w, k)
        return bool((r == (bv.rank(idx_prefix, 1) & mask)).all() and (s_ == (bv.select(sel_prefix, 1) & mask)).all())

    def describe(self):
        return (f"sdslgpu_rank_iv / sdslgpu_select_iv: queries and results as int_vector<{self.w}> fields in pinned host arrays "
                f"({self.w / 8:g} B per query and per result over PCIe instead of 8; chunks of 2^23 queries: H2D, unpack, kernels, pack, D2H overlapped)")


def binned_wanted(index_bytes, n, select=False):
    """what SDSLGPU_ORDER_AUTO resolves to for a batch of n queries on an index of index_bytes (sdslgpu_auto_is_binned)"""
    return bool(lib().sdslgpu_auto_is_binned(int(index_bytes), int(n), int(bool(select))))


class BitVector(_Handle):
    """bit_vector + rank_support_v<b> + select_support_mcl<b> on the device.

    ``words``: ceil(nbits/64) uint64 (numpy, or a torch CUDA int64 tensor already resident in HBM).
    """

    def __init__(self, words, nbits, device=0, flags=F_DEFAULT):
        super().__init__()
        nbits = int(nbits)
        if _is_torch(words):
            p, n, keep = words.data_ptr(), words.numel(), words
        else:
            keep = np.ascontiguousarray(words, dtype=np.uint64)
            p, n = keep.ctypes.data, keep.size
        assert n >= (nbits + 63) // 64, "words too short for nbits"
        _check(lib().sdslgpu_bv_create(p if n else None, nbits, device, flags, C.byref(self._h)))
        self.nbits = nbits
        self.flags = flags



class _WaveletTreeOps:
    """wt.rank(i, c) / wt.select(i, c) / wt[i] / inverse_select(i) in batch form (SURVEY §8(b) wavelet-tree concept)"""

    _sym_dtype = np.uint8

    @property
    def sigma(self):
        v = C.c_uint64()
        _check(lib().sdslgpu_wt_sigma(self._h, C.byref(v)))
        return v.value

    def wt_rank(self, i, c, out=None, stream=None):
        p, n, k1, _ = _in_ptr(i)
        pc, nc, k2, _ = _in_ptr(c, self._sym_dtype)
        assert n == nc
        po, o, _k = _out_like(i, n, out)
        _check(lib().sdslgpu_wt_rank(self._h, p, pc, n, po, _stream_ptr(stream, i)))
        return o

    def wt_select(self, i, c, out=None, stream=None):
        p, n, k1, _ = _in_ptr(i)
        pc, nc, k2, _ = _in_ptr(c, self._sym_dtype)
        assert n == nc
        po, o, _k = _out_like(i, n, out)
        _check(lib().sdslgpu_wt_select(self._h, p, pc, n, po, _stream_ptr(stream, i)))
        return o

    def wt_access(self, i, stream=None):
        p, n, k1, _ = _in_ptr(i)
        po, o, _k = _out_like(i, n)
        _check(lib().sdslgpu_wt_access(self._h, p, n, po, None, _stream_ptr(stream, i)))
        return o

    def inverse_select(self, i, stream=None):
        """-> (rank(i, wt[i]), wt[i])  (wt_pc.hpp:411-430)"""
        p, n, k1, _ = _in_ptr(i)
        ps, sym, _k = _out_like(i, n)
        pr, rnk, _k2 = _out_like(i, n)
        _check(lib().sdslgpu_wt_access(self._h, p, n, ps, pr, _stream_ptr(stream, i)))
        return rnk, sym


class WtHuff(_Handle, _WaveletTreeOps):
    """wt_huff<> over a byte text (host bytes / numpy uint8)."""

    def __init__(self, text, device=0, flags=F_DEFAULT):
        super().__init__()
        t = np.frombuffer(text, dtype=np.uint8) if isinstance(text, (bytes, bytearray)) else np.ascontiguousarray(text, dtype=np.uint8)
        self._text_keep = t
        _check(lib().sdslgpu_wt_huff_create(t.ctypes.data if len(t) else None, len(t), device, flags, C.byref(self._h)))

    rank = _WaveletTreeOps.wt_rank
    select = _WaveletTreeOps.wt_select
    access = _WaveletTreeOps.wt_access


_SIGNATURES += [
    ("sdslgpu_csa_create", C.c_int, [vp, C.c_uint64, C.c_int, C.c_uint32, C.POINTER(vp)]),
    ("sdslgpu_csa_create_ex", C.c_int, [vp, C.c_uint64, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(vp)]),
    ("sdslgpu_fm_count", C.c_int, [vp, vp, vp, C.c_uint64, vp, vp, vp]),
    ("sdslgpu_csa_alphabet", C.c_int, [vp, vp, vp, vp, C.POINTER(C.c_uint32)]),
    ("sdslgpu_fm_sa", C.c_int, [vp, vp, C.c_uint64, vp, vp]),
    ("sdslgpu_fm_locate", C.c_int, [vp, vp, vp, C.c_uint64, vp, vp, C.c_uint64, u64p, vp]),
    ("sdslgpu_fm_extract", C.c_int, [vp, vp, vp, C.c_uint64, vp, vp, vp]),
]


def csr_patterns(pats):
    """list of bytes -> (uint8 concatenation, uint64 offsets[n+1])"""
    off = np.zeros(len(pats) + 1, dtype=np.uint64)
    if len(pats):
        off[1:] = np.cumsum([len(p) for p in pats], dtype=np.uint64)
    flat = np.frombuffer(b"".join(pats), dtype=np.uint8).copy()
    if len(flat) == 0:
        flat = np.zeros(1, np.uint8)
    return flat, off


class CsaWt(_Handle, _WaveletTreeOps):
    """csa_wt<wt_huff<>> over a zero-free byte text; count / locate / SA access in batch form
    (sdsl::count, sdsl::locate, csa[i] — suffix_array_algorithm.hpp:463-471, 534-550; csa_wt.hpp:363-381)."""

    def __init__(self, text, device=0, flags=F_DEFAULT, sa_dens=0, isa_dens=0):
        """sa_dens / isa_dens are csa_wt's t_dens / t_inv_dens (csa_wt.hpp:50-51); 0 = the defaults 32 / 64"""
        super().__init__()
        t = np.frombuffer(text, dtype=np.uint8) if isinstance(text, (bytes, bytearray)) else np.ascontiguousarray(text, dtype=np.uint8)
        _check(lib().sdslgpu_csa_create_ex(t.ctypes.data if len(t) else None, len(t), device, flags, sa_dens, isa_dens, C.byref(self._h)))

    # csa.bwt.rank(i, c) etc.
    bwt_rank = _WaveletTreeOps.wt_rank

    def alphabet(self):
        """-> (C uint64[257], char2comp uint8[256], comp2char uint8[256], sigma): csa.C / .char2comp / .comp2char / .sigma"""
        Cs, c2c, cc2 = np.zeros(257, np.uint64), np.zeros(256, np.uint8), np.zeros(256, np.uint8)
        sg = C.c_uint32()
        _check(lib().sdslgpu_csa_alphabet(self._h, Cs.ctypes.data, c2c.ctypes.data, cc2.ctypes.data, C.byref(sg)))
        return Cs, c2c, cc2, sg.value

    def count(self, flat, off, want_l=False, stream=None):
        """flat: uint8 pattern bytes, off: uint64[n+1] (numpy or torch CUDA tensors)"""
        pf, nf, k1, _ = _in_ptr(flat, np.uint8)
        po, no, k2, _ = _in_ptr(off)
        n = no - 1
        pc, cnt, _k = _out_like(off, n)
        pl, l, _k2 = _out_like(off, n) if want_l else (None, None, None)
        _check(lib().sdslgpu_fm_count(self._h, pf, po, n, pc, pl, _stream_ptr(stream, off)))
        return (cnt, l) if want_l else cnt

    def sa(self, i, stream=None):
        p, n, k1, _ = _in_ptr(i)
        po, o, _k = _out_like(i, n)
        _check(lib().sdslgpu_fm_sa(self._h, p, n, po, _stream_ptr(stream, i)))
        return o

    def extract(self, begin, end, stream=None):
        """text[begin[k] .. end[k]] (inclusive) for every k -> (offsets uint64[n+1], uint8 bytes); host arrays"""
        b = np.ascontiguousarray(begin, dtype=np.uint64)
        e = np.ascontiguousarray(end, dtype=np.uint64)
        off = np.zeros(len(b) + 1, dtype=np.uint64)
        off[1:] = np.cumsum(e - b + np.uint64(1), dtype=np.uint64)
        out = np.zeros(max(int(off[-1]), 1), dtype=np.uint8)
        _check(lib().sdslgpu_fm_extract(self._h, b.ctypes.data, e.ctypes.data, len(b), off.ctypes.data, out.ctypes.data, _stream_ptr(stream, b)))
        return off, out[: int(off[-1])]

    def locate(self, flat, off, stream=None):
        """-> (occ_off uint64[n+1], occ uint64[total]) with each pattern's occurrences in suffix-array order"""
        pf, nf, k1, _ = _in_ptr(flat, np.uint8)
        po, no, k2, _ = _in_ptr(off)
        n = no - 1
        poo, occ_off, _k = _out_like(off, n + 1)
        total = C.c_uint64()
        sp = _stream_ptr(stream, off)
        # one pass when the guessed capacity suffices (one backward search per pattern); otherwise the call
        # reports the exact total and a second pass fills a buffer of that size
        cap = max(2 * n, 1 << 16)
        pocc, occ, _k2 = _out_like(off, cap)
        st = lib().sdslgpu_fm_locate(self._h, pf, po, n, poo, pocc, cap, C.byref(total), sp)
        if st == EINVAL and total.value > cap:
            pocc, occ, _k2 = _out_like(off, total.value)
            st = lib().sdslgpu_fm_locate(self._h, pf, po, n, poo, pocc, total.value, C.byref(total), sp)
        _check(st)
        return occ_off, occ[: total.value]


_SIGNATURES += [
    ("sdslgpu_rrr63_create", C.c_int, [vp, C.c_uint64, C.c_int, C.c_uint32, C.POINTER(vp)]),
    ("sdslgpu_sd_create", C.c_int, [vp, C.c_uint64, C.c_int, C.c_uint32, C.POINTER(vp)]),
    ("sdslgpu_serialize", C.c_int, [vp, C.c_int, vp, C.c_uint64, u64p]),
]


class _CompressedBitVector(_Handle):
    _create = None

    def __init__(self, words, nbits, device=0, flags=F_DEFAULT):
        super().__init__()
        nbits = int(nbits)
        if _is_torch(words):
            p, n, keep = words.data_ptr(), words.numel(), words
        else:
            keep = np.ascontiguousarray(words, dtype=np.uint64)
            p, n = keep.ctypes.data, keep.size
        assert n >= (nbits + 63) // 64, "words too short for nbits"
        _check(getattr(lib(), self._create)(p if n else None, nbits, device, flags, C.byref(self._h)))
        self.nbits = nbits

class RrrVector(_CompressedBitVector):
    """rrr_vector<63> + rank_support_rrr + select_support_rrr, encoded on the device"""

    _create = "sdslgpu_rrr63_create"


class SdVector(_CompressedBitVector):
    """sd_vector<> + rank_support_sd + select_support_sd, built on the device"""

    _create = "sdslgpu_sd_create"


_SIGNATURES += [
    ("sdslgpu_wt_int_create", C.c_int, [vp, C.c_uint64, C.c_int, C.c_uint32, C.POINTER(vp)]),
]


_SIGNATURES += [
    ("sdslgpu_load_sdsl", C.c_int, [vp, C.c_uint64, C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.POINTER(vp)]),
    ("sdslgpu_load_sdsl_ex", C.c_int, [vp, C.c_uint64, C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, u64p, C.POINTER(vp)]),
]


def load_sdsl(blob, kind, device=0, flags=F_DEFAULT, param=0, isa_dens=0, want_consumed=False):
    """Ingest bytes written by the reference's serialize()/store_to_file -> a handle object of the right class.
    param = t_dens of a CSA, isa_dens = its t_inv_dens (0: inferred); want_consumed -> (object, bytes of the blob used)."""
    cls = {KIND_BV: BitVector, KIND_RRR63: RrrVector, KIND_SD: SdVector, KIND_WT_HUFF: WtHuff, KIND_WT_INT: WtInt, KIND_CSA_WT: CsaWt}[kind]
    obj = cls.__new__(cls)
    _Handle.__init__(obj)
    buf = np.frombuffer(blob, dtype=np.uint8)
    used = C.c_uint64(0)
    _check(lib().sdslgpu_load_sdsl_ex(buf.ctypes.data, len(buf), kind, device, flags, param, isa_dens, C.byref(used), C.byref(obj._h)))
    obj.flags = flags
    obj.nbits = obj.size
    return (obj, used.value) if want_consumed else obj


class WtInt(_Handle, _WaveletTreeOps):
    """wt_int<> over a sequence of unsigned integers (host numpy uint64)"""

    _sym_dtype = np.uint64

    def __init__(self, seq, device=0, flags=F_DEFAULT):
        super().__init__()
        s = np.ascontiguousarray(seq, dtype=np.uint64)
        _check(lib().sdslgpu_wt_int_create(s.ctypes.data if len(s) else None, len(s), device, flags, C.byref(self._h)))

    rank = _WaveletTreeOps.wt_rank
    select = _WaveletTreeOps.wt_select
    access = _WaveletTreeOps.wt_access


# ------------------------------------------------------------------------------------------------------
# multi-GPU groups (include/sdslgpu.h "multi-GPU groups"; csrc/group.cu)
# ------------------------------------------------------------------------------------------------------
GATHER_NONE, GATHER_NCCL, GATHER_FUSED, GATHER_AUTO, GATHER_PACKED = 0, 1, 2, 3, 4
UNIQUE_ID_BYTES = 128
vpp = C.POINTER(vp)

_SIGNATURES += [
    ("sdslgpu_group_unique_id", C.c_int, [vp]),
    ("sdslgpu_group_create_rank", C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]),
    ("sdslgpu_group_create", C.c_int, [C.POINTER(C.c_int), C.c_int, C.POINTER(vp)]),
    ("sdslgpu_group_free", C.c_int, [vp]),
    ("sdslgpu_group_info", C.c_int, [vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    ("sdslgpu_group_alloc", C.c_int, [vp, C.c_uint64, vpp]),
    ("sdslgpu_group_release", C.c_int, [vp, vpp]),
    ("sdslgpu_group_replicate", C.c_int, [vp, vp, C.c_int, vpp]),
    ("sdslgpu_group_rank", C.c_int, [vp, vpp, C.c_int, vpp, C.c_uint64, vpp, C.c_int, vpp]),
    ("sdslgpu_group_select", C.c_int, [vp, vpp, C.c_int, vpp, C.c_uint64, vpp, C.c_int, vpp]),
    ("sdslgpu_group_wt_rank", C.c_int, [vp, vpp, vpp, vpp, C.c_uint64, vpp, C.c_int, vpp]),
    ("sdslgpu_group_fm_count", C.c_int, [vp, vpp, vpp, vpp, C.c_uint64, vpp, C.c_int, vpp]),
]


def group_unique_id():
    """the 128 bytes rank 0 hands to every other rank before Group.create_rank (ncclGetUniqueId)"""
    buf = (C.c_uint8 * UNIQUE_ID_BYTES)()
    _check(lib().sdslgpu_group_unique_id(buf))
    return bytes(buf)


class SymmetricBuffer:
    """`bytes` of device memory on every local member of a group, mapped by all other members (sdslgpu_group_alloc);
    tensor(k, dtype) views local member k's copy as a torch tensor without copying"""

    def __init__(self, group, nbytes):
        self.group, self.nbytes = group, int(nbytes)
        self.ptrs = (vp * group.nlocal)()
        _check(lib().sdslgpu_group_alloc(group._g, self.nbytes, self.ptrs))

    def tensor(self, k=0, dtype=None):
        import torch

        dtype = dtype or torch.int64
        dev = self.group.devices[k]

        class _Raw:  # the CUDA array interface torch.as_tensor understands
            pass

        raw = _Raw()
        raw.__cuda_array_interface__ = {"shape": (self.nbytes,), "typestr": "|u1", "data": (int(self.ptrs[k]), False), "version": 2}
        t = torch.as_tensor(raw, device=torch.device("cuda", dev))
        self._keep = getattr(self, "_keep", []) + [raw]
        return t.view(dtype)

    def release(self):
        if self.ptrs is not None and self.group._g:
            _check(lib().sdslgpu_group_release(self.group._g, self.ptrs))
        self.ptrs = None


class Group:
    """Replicated index, sharded batch, all-gathered results.

    Group.create(devices)           one process, several devices: every call takes LISTS with one entry per device
    Group.create_rank(id, n, r, d)  one process per GPU (torchrun): lists of length one

    rank / select / wt_rank / fm_count take per-member handle objects and torch CUDA tensors; every member's input holds
    the WHOLE batch, every member's output receives ALL results.  gather = GATHER_NONE / _NCCL / _FUSED / _AUTO."""

    def __init__(self):
        self._g = vp()
        self.devices = []

    @classmethod
    def create(cls, devices):
        g = cls()
        arr = (C.c_int * len(devices))(*[int(d) for d in devices])
        _check(lib().sdslgpu_group_create(arr, len(devices), C.byref(g._g)))
        g.devices = [int(d) for d in devices]
        g._info()
        return g

    @classmethod
    def create_rank(cls, unique_id, nranks, rank, device):
        g = cls()
        assert len(unique_id) == UNIQUE_ID_BYTES
        buf = (C.c_uint8 * UNIQUE_ID_BYTES).from_buffer_copy(unique_id)
        _check(lib().sdslgpu_group_create_rank(buf, int(nranks), int(rank), int(device), C.byref(g._g)))
        g.devices = [int(device)]
        g._info()
        return g

    def _info(self):
        a, b, c, d = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        _check(lib().sdslgpu_group_info(self._g, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        self.nranks, self.nlocal, self.first_rank, self.fused_possible = a.value, b.value, c.value, bool(d.value)

    def close(self):
        if self._g:
            lib().sdslgpu_group_free(self._g)
            self._g = vp()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def alloc(self, nbytes):
        return SymmetricBuffer(self, nbytes)

    def replicate(self, src, root=0):
        """src: a handle object on the process that owns global rank `root` (None elsewhere) -> list of replicas"""
        out = (vp * self.nlocal)()
        _check(lib().sdslgpu_group_replicate(self._g, src._h if src is not None else None, int(root), out))
        res = []
        for k in range(self.nlocal):
            kind = C.c_int()
            h = vp(out[k])
            _check(lib().sdslgpu_kind(h, C.byref(kind)))
            cls = {KIND_BV: BitVector, KIND_RRR63: RrrVector, KIND_SD: SdVector, KIND_WT_HUFF: WtHuff, KIND_WT_INT: WtInt, KIND_CSA_WT: CsaWt}[kind.value]
            obj = cls.__new__(cls)
            _Handle.__init__(obj)
            obj._h = h
            obj.nbits = obj.size
            res.append(obj)
        return res

    def _ptrs(self, xs, what):
        assert len(xs) == self.nlocal, f"{what}: one entry per local member ({self.nlocal})"
        return (vp * self.nlocal)(*[int(x.data_ptr()) if _is_torch(x) else int(x) for x in xs])

    def _handles(self, hs):
        assert len(hs) == self.nlocal
        return (vp * self.nlocal)(*[h._h.value for h in hs])

    def _streams(self, streams):
        if streams is None:
            return None
        return (vp * self.nlocal)(*[int(getattr(s, "cuda_stream", s)) for s in streams])

    def rank(self, hs, b, idx, out, gather=GATHER_AUTO, streams=None):
        _check(lib().sdslgpu_group_rank(self._g, self._handles(hs), int(b), self._ptrs(idx, "idx"), int(idx[0].numel()), self._ptrs(out, "out"),
                                        int(gather), self._streams(streams)))
        return out

    def select(self, hs, b, i, out, gather=GATHER_AUTO, streams=None):
        _check(lib().sdslgpu_group_select(self._g, self._handles(hs), int(b), self._ptrs(i, "i"), int(i[0].numel()), self._ptrs(out, "out"),
                                          int(gather), self._streams(streams)))
        return out

    def wt_rank(self, hs, i, c, out, gather=GATHER_AUTO, streams=None):
        _check(lib().sdslgpu_group_wt_rank(self._g, self._handles(hs), self._ptrs(i, "i"), self._ptrs(c, "c"), int(i[0].numel()),
                                           self._ptrs(out, "out"), int(gather), self._streams(streams)))
        return out

    def fm_count(self, hs, flat, off, out, gather=GATHER_AUTO, streams=None):
        _check(lib().sdslgpu_group_fm_count(self._g, self._handles(hs), self._ptrs(flat, "pats"), self._ptrs(off, "pat_off"), int(off[0].numel()) - 1,
                                            self._ptrs(out, "cnt_out"), int(gather), self._streams(streams)))
        return out
