/* sdslgpu.h — C ABI of the B200-native rank/select + wavelet-tree + FM-index query engine.
 *
 * This is the drop-in boundary for the hot path named in BASELINE.json.  The
 * reference (xxsds/sdsl-lite 3.0.5) has no FFI of its own — it is a header-only template library
 * whose composition points are duck-typed concepts — so every entry point below cites the
 * reference member function(s) it replaces (paths relative to /root/reference/include/sdsl/).
 * The SDSL-shaped C++ classes on top of this ABI live in sdsl-lite_b200/include/sdsl_b200.hpp;
 * the binding a maintainer of the reference would add is shown in INTEGRATION.md.
 *
 * Conventions
 *  - plain C types only; handles are opaque; no function throws or aborts;
 *  - every function returns a status: SDSLGPU_OK (0) or a negative SDSLGPU_E* code;
 *    sdslgpu_last_error() returns a thread-local message for the last failure;
 *  - batch calls take n inputs and write n outputs in the same order;
 *  - input/output pointers may be HOST or DEVICE pointers (detected per pointer with
 *    cudaPointerGetAttributes).  Device pointers must live on the handle's device; the call is
 *    then asynchronous on `stream`.  Host pointers are streamed through per-handle device chunk
 *    buffers (3 slots x 2^22 queries, allocated on first use; H2D, kernel and D2H overlapped —
 *    pin the host arrays with cudaHostRegister / cudaHostAlloc to get full PCIe speed) and the
 *    call returns after the results are in `out`;
 *  - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream);
 *  - the SDSL preconditions that are undefined behaviour in the reference (rank idx > size,
 *    select i == 0 or i > #args) are DEFINED here: the result for that query is
 *    SDSLGPU_NPOS (all ones) and the call still returns SDSLGPU_OK.  One in-band result of the
 *    reference is kept instead: select on a KIND_RRR63 handle with i > #args returns size(), as
 *    select_support_rrr does (rrr_vector.hpp:641-642, 686-689); i == 0 is SDSLGPU_NPOS there too;
 *  - there is NO CPU fallback: without a CUDA device every create call fails with SDSLGPU_ECUDA.
 */
#ifndef SDSLGPU_H
#define SDSLGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDSLGPU_OK 0
#define SDSLGPU_EINVAL (-1)   /* bad argument (null pointer, unknown kind/pattern, malformed blob) */
#define SDSLGPU_ENOMEM (-2)   /* host or device allocation failed */
#define SDSLGPU_ECUDA (-3)    /* CUDA runtime error (no device, launch failure, ...) */
#define SDSLGPU_ENOTSUP (-4)  /* operation not supported by this handle kind */
#define SDSLGPU_NPOS UINT64_MAX

/* handle kinds (sdslgpu_kind) */
#define SDSLGPU_KIND_BV 1       /* bit_vector + rank_support_v<b> + select_support_mcl<b>           */
#define SDSLGPU_KIND_RRR63 2    /* rrr_vector<63> + rank_support_rrr + select_support_rrr            */
#define SDSLGPU_KIND_SD 3       /* sd_vector<> + rank_support_sd + select_support_sd                 */
#define SDSLGPU_KIND_WT_HUFF 4  /* wt_huff<>                                                         */
#define SDSLGPU_KIND_WT_INT 5   /* wt_int<>                                                          */
#define SDSLGPU_KIND_CSA_WT 6   /* csa_wt<wt_huff<>, 32, 64, sa_order_sa_sampling<>, isa_sampling<>> */

/* creation flags */
#define SDSLGPU_F_DEFAULT 0u
#define SDSLGPU_F_SDSL_LAYOUT 1u /* additionally keep SDSL's own device layout (plain words + the
                                    rank_support_v m_basic_block table) and answer rank from it; the
                                    default is the sector-interleaved B200 layout */
#define SDSLGPU_F_NO_SELECT 2u   /* skip building the select samples (rank-only handle) */
#define SDSLGPU_F_RRR_BV 4u      /* wavelet trees / CSA: store the tree's bit vector as rrr_vector<63>, i.e.
                                    wt_huff<rrr_vector<63>> and csa_wt<wt_huff<rrr_vector<63>>> (H0-compressed) */

#define SDSLGPU_F_V5_SCAN 16u    /* sdslgpu_load_sdsl of a wavelet tree / CSA: the blob is the reference's
                                    wt_huff<bit_vector, rank_support_v5<>, select_support_scan<>, select_support_scan<0>>
                                    form (its own count benchmark's FM_HUFF index, benchmark/indexing_count/
                                    index.config:8): a rank_support_v5 table and no select data follow m_bv */
#define SDSLGPU_F_COMPACT 8u     /* CSA: keep the wavelet tree of the BWT as the only occurrence structure
                                    (1.14 bytes/symbol; a backward-search step is one sector gather per Huffman
                                    level).  By default a CSA additionally holds 32 one-hot sector-block bitmaps
                                    and the BWT (5.6 bytes/symbol) so that a step is 2 gathers and
                                    an LF step 3; results are identical.  Implied by SDSLGPU_F_RRR_BV.
                                    Bit vectors (KIND_BV): never build select sectors (see sdslgpu_select). */

/* bit patterns of sdslgpu_rank / sdslgpu_select / sdslgpu_arg_count on KIND_BV handles: the reference's
 * <t_b, t_pat_len> template arguments (rank_support_v.hpp:46, select_support_mcl.hpp:54).  An occurrence of a
 * two-bit pattern "xy" is counted / reported at the position of y, its second bit (x = the bit before it), as
 * in the reference (select_support.hpp:233-238). */
#define SDSLGPU_PAT_0 0
#define SDSLGPU_PAT_1 1
#define SDSLGPU_PAT_10 2 /* <10,2>: a 1 followed by a 0 */
#define SDSLGPU_PAT_01 3 /* <01,2> */
#define SDSLGPU_PAT_00 4 /* <00,2> */
#define SDSLGPU_PAT_11 5 /* <11,2> */

typedef struct sdslgpu_handle sdslgpu_handle;

/* ---- library ------------------------------------------------------------------------------- */
const char *sdslgpu_version(void);
const char *sdslgpu_last_error(void);
int sdslgpu_device_count(int *count);

/* ---- creation / destruction ---------------------------------------------------------------- */

/* Replaces: bit_vector(words) + rank_support_v<1>/<0> ctor (rank_support_v.hpp:72-122) +
 * select_support_mcl<1>/<0> ctor (select_support_mcl.hpp:121-128).  `words` holds ceil(nbits/64)
 * little-endian 64-bit words, bit i = (words[i>>6] >> (i&63)) & 1 (int_vector.hpp:1900-1904); bits
 * past nbits in the last word are ignored.  All supports are built on the device. */
int sdslgpu_bv_create(const uint64_t *words, uint64_t nbits, int device, uint32_t flags, sdslgpu_handle **out);

/* Replaces rrr_vector<63>(bit_vector) (rrr_vector.hpp:158-270): classes, enumerative offsets (bin_to_nr,
 * rrr_helper.hpp:346-366), per-superblock samples and invert bits are computed ON THE DEVICE and are
 * bit-identical to the reference's (sdslgpu_serialize returns the reference's byte format). */
int sdslgpu_rrr63_create(const uint64_t *words, uint64_t nbits, int device, uint32_t flags, sdslgpu_handle **out);

/* Replaces sd_vector<>(bit_vector) (sd_vector.hpp:218-257): low / high parts are scattered on the device
 * (bit-identical to m_low / m_high); select_support_mcl<1>/<0> over `high` (sd_vector.hpp:162-163) are
 * served by the engine's own rank blocks + select samples. */
int sdslgpu_sd_create(const uint64_t *words, uint64_t nbits, int device, uint32_t flags, sdslgpu_handle **out);

int sdslgpu_free(sdslgpu_handle *h);
int sdslgpu_kind(const sdslgpu_handle *h, int *kind);
/* size(): number of bits (bit vectors) / symbols (wavelet trees) / text length + 1 (csa) */
int sdslgpu_size(const sdslgpu_handle *h, uint64_t *size);
/* number of b-bits (occurrences of pattern b) in a bit-vector handle (= rank_b(size)); select domain is 1..arg_count */
int sdslgpu_arg_count(const sdslgpu_handle *h, int b, uint64_t *count);
/* bytes of device memory held by the handle's index structures, including what queries derive on first use (pattern
 * images, select sectors); NOT counted: the chunk buffers of the host-pointer path (3 slots x 2^22 queries x 32 bytes =
 * 0.4 GB, allocated by the handle's first call with host pointers; the int_vector<w> path has its own of the same order) */
int sdslgpu_device_bytes(const sdslgpu_handle *h, uint64_t *bytes);

/* ---- batched queries on bit vectors --------------------------------------------------------- */

/* out[k] = number of b-bits in [0, idx[k]),  0 <= idx[k] <= size.
 * Replaces rank_support_v<b,1>::rank (rank_support_v.hpp:129-139),
 *          rank_support_rrr<b,63>::rank (rrr_vector.hpp:503-544),
 *          rank_support_sd<b>::rank (sd_vector.hpp:553-575).   b in {0,1}.
 * On KIND_BV handles b may also be SDSLGPU_PAT_10 / _01 / _00 / _11: rank_support_v<10|01|00|11, 2>::rank
 * (rank_support.hpp:161-284).  The first such call builds that pattern's indicator vector on the device
 * (+ 1.14 bits/bit for rank, + samples for select) and later calls reuse it. */
int sdslgpu_rank(const sdslgpu_handle *h, int b, const uint64_t *idx, uint64_t n, uint64_t *out, void *stream);

/* out[k] = position of the i[k]-th b-bit, 1 <= i[k] <= arg_count(b).
 * Replaces select_support_mcl<b,1>::select (select_support_mcl.hpp:384-439),
 *          select_support_rrr<b,63>::select (rrr_vector.hpp:639-726),
 *          select_support_sd<b>::select (sd_vector.hpp:621-664);
 *          on KIND_BV also select_support_mcl<10|01|00|11, 2>::select (select_support.hpp:204-405). */
int sdslgpu_select(const sdslgpu_handle *h, int b, const uint64_t *i, uint64_t n, uint64_t *out, void *stream);
/* Memory note (KIND_BV, b in {0,1}).  The first select batch that SDSLGPU_ORDER_AUTO / _BINNED runs through the binned
 * pipeline builds "select sectors" for that b (csrc/bv_device.cuh): one 32-byte record per S b-bits holding the position
 * of the first of them and the 193 - 224 bits of the vector that follow it, so that a query is ONE gather (the sampled
 * select it replaces: a sample pair, a block, and a neighbouring block for 5 % of the queries).  S follows from the
 * density (81 at 1/2); the records cost 32/S bytes per b-bit — 1.7 GB for the ones of a 2^33-bit vector of density 1/2
 * whose rank/select image is 1.3 GB — and are counted by sdslgpu_device_bytes once built.  They are NOT built for
 * handles created with SDSLGPU_F_COMPACT, densities under ~8 %, vectors beyond 2^36 bits, or when less than the records
 * + the batch's scratch + 1 GiB of device memory is free; the sampled select then keeps serving.  Results are identical either way.  The
 * build (a few ms) synchronises the device once: do not make that first call inside a stream capture. */

/* The same two calls with queries and results in the reference's own compact container: int_vector<w>.  Field k of
 * width w occupies bits [k*w, (k+1)*w) of the word array, LSB first (int_vector.hpp:1900-1904 for w = 1, get_int /
 * bits::read_int bits.hpp:777-790 in general) — i.e. `words` is int_vector<>::data() of a vector the caller filled or
 * bit-compressed (util::bit_compress, util.hpp:502-516).  Positions in a 2^33-bit vector need 34 bits, not 64: the
 * batch crosses PCIe at 4.25 bytes per query and per result instead of 8.  idx_words holds ceil(n*idx_width/64)
 * words, out_words receives ceil(n*out_width/64) words; a result is truncated to out_width bits (choose
 * out_width >= bits::hi(size)+1; SDSLGPU_NPOS becomes the all-ones field).  Both arrays on the host (chunks of 2^23
 * queries: H2D of the packed chunk, unpack / kernels / pack on the device, D2H of the packed results, overlapped) or
 * both on the device (asynchronous on `stream`).  Values are identical to sdslgpu_rank / sdslgpu_select. */
int sdslgpu_rank_iv(const sdslgpu_handle *h, int b, const uint64_t *idx_words, uint32_t idx_width, uint64_t n,
                    uint64_t *out_words, uint32_t out_width, void *stream);
int sdslgpu_select_iv(const sdslgpu_handle *h, int b, const uint64_t *i_words, uint32_t i_width, uint64_t n,
                      uint64_t *out_words, uint32_t out_width, void *stream);

/* Order of work inside one sdslgpu_rank / sdslgpu_select call on a KIND_BV handle; never changes a result.
 *   SDSLGPU_ORDER_DIRECT  one thread per query in the caller's order: one random DRAM gather per query
 *   SDSLGPU_ORDER_BINNED  the batch is counting-sorted tile by tile into 16 - 24 MB chunks of the index, answered chunk
 *                         by chunk (L2-resident gathers) and un-sorted again; needs ~14 bytes of stream-ordered
 *                         scratch per query (cudaMallocAsync on the call's stream)
 *   SDSLGPU_ORDER_AUTO    (default) BINNED when the index is larger than the L2 (>= 192 MB) and the batch is dense
 *                         enough for queries to share cache lines (>= 2^21 queries and >= 1 query per 128 bytes of
 *                         index for rank and for select through select sectors, per 192 bytes for the sampled
 *                         select — the measured break-even points), else DIRECT
 * The reference has no counterpart (its queries are scalar calls, rank_support_v.hpp:129-139). */
#define SDSLGPU_ORDER_AUTO 0
#define SDSLGPU_ORDER_DIRECT 1
#define SDSLGPU_ORDER_BINNED 2
int sdslgpu_set_batch_order(sdslgpu_handle *h, int order);
/* 1 if SDSLGPU_ORDER_AUTO runs a batch of n rank (select != 0: select) queries on an index of index_bytes through the
 * binned pipeline, else 0 (for callers that want to report or plan around the choice; no handle, no device needed) */
int sdslgpu_auto_is_binned(uint64_t index_bytes, uint64_t n, int select);

/* out[k] = bit idx[k] (0/1), 0 <= idx[k] < size.  Replaces operator[] of bit_vector
 * (int_vector.hpp:1900-1904), rrr_vector (rrr_vector.hpp:276-298), sd_vector (sd_vector.hpp:328-349). */
int sdslgpu_access(const sdslgpu_handle *h, const uint64_t *idx, uint64_t n, uint64_t *out, void *stream);

/* ---- wavelet trees ---------------------------------------------------------------------------- */

/* Replaces wt_huff<>(text) = wt_pc<huff_shape,...> ctor (wt_pc.hpp:194-248, wt_huff.hpp:82-115,
 * wt_helper.hpp:230-327).  `text` is a HOST buffer of n bytes (any byte values).  The Huffman shape
 * (<= 511 nodes) is computed on the host with the reference's tie-breaking; the concatenated bit vector m_bv is
 * built on the device — one stable radix pass per tree depth over the device-resident text (wt_build.cu) — and is
 * bit-identical to the reference's, as are the node table and the serialised tree (sdslgpu_serialize); rank / select
 * structures over it are built on the device.  (SDSLGPU_HOST_WT=1, or no device memory for the 6 bytes/symbol of
 * sort scratch: the multi-threaded host fill produces the same bits.) */
int sdslgpu_wt_huff_create(const uint8_t *text, uint64_t n, int device, uint32_t flags, sdslgpu_handle **out);

/* Replaces wt_int<>(seq) (wt_int.hpp:160-260): `seq` is a HOST array of n integers; the level bit vector
 * m_tree (max_level = hi(max)+1 levels of n bits) is laid out exactly like the reference's; like wt_huff it is built
 * on the device (level k = the sequence stably radix-sorted by its top k bits). */
int sdslgpu_wt_int_create(const uint64_t *seq, uint64_t n, int device, uint32_t flags, sdslgpu_handle **out);

/* number of distinct symbols (wt.sigma, wt_pc.hpp:177) */
int sdslgpu_wt_sigma(const sdslgpu_handle *h, uint64_t *sigma);

/* out[k] = occurrences of symbol c[k] in [0, i[k]),  0 <= i[k] <= size.  A symbol that does not occur
 * gives 0.  `c` points to uint8_t symbols for byte trees (KIND_WT_HUFF, KIND_CSA_WT) and to uint64_t
 * symbols for KIND_WT_INT.  Replaces wt_pc::rank (wt_pc.hpp:371-399), wt_int::rank (wt_int.hpp:379-409). */
int sdslgpu_wt_rank(const sdslgpu_handle *h, const uint64_t *i, const void *c, uint64_t n, uint64_t *out, void *stream);

/* out[k] = position of the i[k]-th occurrence of c[k], 1 <= i[k] <= rank(size, c[k]).  A symbol that does
 * not occur gives size() (as the reference, wt_pc.hpp:447-450); i beyond the occurrences gives
 * SDSLGPU_NPOS.  Replaces wt_pc::select (wt_pc.hpp:443-474), wt_int::select (wt_int.hpp:456-507). */
int sdslgpu_wt_select(const sdslgpu_handle *h, const uint64_t *i, const void *c, uint64_t n, uint64_t *out, void *stream);

/* sym_out[k] = wt[i[k]]; if rank_out != NULL also rank_out[k] = rank(i[k], wt[i[k]]) (inverse_select).
 * Replaces wt_pc::operator[] / inverse_select (wt_pc.hpp:336-357, 411-430), wt_int (wt_int.hpp:340-367,
 * 418-445). */
int sdslgpu_wt_access(const sdslgpu_handle *h, const uint64_t *i, uint64_t n, uint64_t *sym_out, uint64_t *rank_out, void *stream);

/* ---- FM-index ---------------------------------------------------------------------------------- */

/* Replaces construct(csa_wt<wt_huff<>>, text) (construct.hpp:127-193, csa_wt.hpp:323-355): `text` is a HOST
 * buffer of n zero-free bytes (a zero byte gives SDSLGPU_EINVAL, as the reference throws, construct.hpp:34-46);
 * the 0 sentinel is appended, so size() == n + 1.  Suffix array, BWT and the SA / ISA samples are computed on the
 * device (prefix doubling; the host SA-IS builder takes over when the text does not fit 32-bit suffix indices or
 * SDSLGPU_HOST_SA=1); the wavelet tree of the BWT gets its rank/select structures on the device.  With
 * SDSLGPU_F_RRR_BV the tree's bit vector is rrr_vector<63>.  A CSA handle also answers sdslgpu_wt_* (that is
 * csa.bwt / csa.wavelet_tree).
 * sdslgpu_csa_create_ex takes the reference's two template parameters (csa_wt.hpp:50-51): sa_dens = t_dens, every
 * sa_dens-th suffix-array entry is kept (csa_sampling_strategy.hpp:98-115), isa_dens = t_inv_dens; 0 means the
 * defaults 32 / 64, which is what sdslgpu_csa_create uses.  Densities change speed and memory, never a result:
 * locate walks on average (sa_dens - 1) / 2 LF steps per occurrence, so with HBM to spare a small sa_dens
 * (4 ... 8) is the B200-side choice. */
int sdslgpu_csa_create_ex(const uint8_t *text, uint64_t n, int device, uint32_t flags, uint32_t sa_dens, uint32_t isa_dens,
                          sdslgpu_handle **out);
int sdslgpu_csa_create(const uint8_t *text, uint64_t n, int device, uint32_t flags, sdslgpu_handle **out);

/* The byte_alphabet of a CSA (csa_alphabet_strategy.hpp:136-212), i.e. the public members csa.C, csa.char2comp,
 * csa.comp2char, csa.sigma (csa_wt.hpp:117-120) that backward_search is written against: C[0..256] (entries past sigma
 * are 0), char2comp[256], comp2char[256] (entries past sigma are 0), *sigma.  Any pointer may be NULL. */
int sdslgpu_csa_alphabet(const sdslgpu_handle *h, uint64_t *C257, uint8_t *char2comp256, uint8_t *comp2char256, uint32_t *sigma);

/* Patterns in CSR form: pattern k = pats[pat_off[k] .. pat_off[k+1]).
 * cnt_out[k] = number of occurrences; if l_out != NULL, l_out[k] = left end of the suffix-array interval
 * (meaningful when cnt_out[k] > 0).  The empty pattern matches size() times; a pattern longer than size()
 * or containing a byte that is not in the text matches 0 times.
 * Replaces sdsl::count / backward_search (suffix_array_algorithm.hpp:166-248, 463-471). */
int sdslgpu_fm_count(const sdslgpu_handle *h, const uint8_t *pats, const uint64_t *pat_off, uint64_t n,
                     uint64_t *cnt_out, uint64_t *l_out, void *stream);

/* out[k] = SA[i[k]] (csa[i], csa_wt.hpp:363-381), 0 <= i[k] < size(). */
int sdslgpu_fm_sa(const sdslgpu_handle *h, const uint64_t *i, uint64_t n, uint64_t *out, void *stream);

/* All occurrences, in SUFFIX-ARRAY order per pattern like sdsl::locate (suffix_array_algorithm.hpp:534-550):
 * occ_off_out[0..n] receives the exclusive prefix sums of the counts, *total_out their sum; if occ_out != NULL
 * it must hold occ_cap >= *total_out entries and receives occ_out[occ_off_out[k] + j] = SA[l_k + j].
 * Call with occ_out == NULL first to size the buffer (or pass a generous occ_cap). */
int sdslgpu_fm_locate(const sdslgpu_handle *h, const uint8_t *pats, const uint64_t *pat_off, uint64_t n,
                      uint64_t *occ_off_out, uint64_t *occ_out, uint64_t occ_cap, uint64_t *total_out, void *stream);

/* out[out_off[k] + j] = text[begin[k] + j] for j = 0 .. end[k] - begin[k]  (end INCLUSIVE, end[k] < size(); the
 * sentinel at position size()-1 reads as 0).  out_off holds n+1 offsets chosen by the caller (normally the
 * exclusive prefix sums of end-begin+1).  Replaces sdsl::extract(csa, begin, end) (suffix_array_algorithm.hpp:590-610)
 * with csa.isa (suffix_array_helper.hpp:519-537) behind it. */
int sdslgpu_fm_extract(const sdslgpu_handle *h, const uint64_t *begin, const uint64_t *end, uint64_t n,
                       const uint64_t *out_off, uint8_t *out, void *stream);

/* ---- construction parity / interchange ------------------------------------------------------ */

/* Copies the SDSL-format serialisation of one component of a KIND_BV handle created with SDSLGPU_F_SDSL_LAYOUT
 * into `buf`, from the words / tables that handle keeps resident, caller's bits past size() included
 * (what: 0 = bit_vector, 1 = rank_support_v<1>, 2 = rank_support_v<0>); *nbytes receives the size
 * needed; buf may be NULL to query it.  The rank tables are built ON THE DEVICE and are byte-identical
 * to rank_support_v::serialize (rank_support_v.hpp:151-158).  (Handles of the default layout: sdslgpu_serialize.) */
int sdslgpu_bv_serialize(const sdslgpu_handle *h, int what, void *buf, uint64_t cap, uint64_t *nbytes);

/* Ingest of the reference's own serialised bytes (store_to_file / serialize(), io.hpp:877-896) for
 * kind = SDSLGPU_KIND_BV        bit_vector                          (int_vector.hpp:1995-2004)
 *        SDSLGPU_KIND_RRR63     rrr_vector<63>                      (rrr_vector.hpp:366-378; taken as is, no re-encoding)
 *        SDSLGPU_KIND_SD        sd_vector<>                         (sd_vector.hpp:426-438)
 *        SDSLGPU_KIND_WT_HUFF   wt_huff<>                           (wt_pc.hpp:713-726, wt_helper.hpp:362-375)
 *        SDSLGPU_KIND_WT_INT    wt_int<>                            (wt_int.hpp:792-805)
 *        SDSLGPU_KIND_CSA_WT    csa_wt<wt_huff<>, t_dens, ...>      (csa_wt.hpp:389-402); param = t_dens (0 -> 32)
 * The rank/select supports stored inside a blob are skipped; this engine builds its own on the device.
 * `blob` is a HOST buffer.  A truncated or inconsistent blob gives SDSLGPU_EINVAL. */
int sdslgpu_load_sdsl(const void *blob, uint64_t nbytes, int kind, int device, uint32_t flags, uint32_t param,
                      sdslgpu_handle **out);
/* The same with both sampling densities of a CSA given explicitly — they are template parameters of the reference
 * (csa_wt.hpp:50-51) and are not stored in the blob: sa_dens = t_dens (0 -> 32), isa_dens = t_inv_dens (0 -> the
 * default 64 if the ISA sample count matches it, else the power of two that does; a count that no candidate
 * reproduces, e.g. an index built with t_inv_dens = 100, is SDSLGPU_EINVAL unless isa_dens is passed).  If
 * consumed != NULL it receives the number of bytes of `blob` the structure occupied, so that a buffer / stream holding
 * several serialised structures can be read one after the other like the reference's load(std::istream&) does. */
int sdslgpu_load_sdsl_ex(const void *blob, uint64_t nbytes, int kind, int device, uint32_t flags, uint32_t sa_dens,
                         uint32_t isa_dens, uint64_t *consumed, sdslgpu_handle **out);

/* Egress: the bytes the reference's serialize() / store_to_file (io.hpp:877-896) writes for the same input, so that
 * an index built here can be stored and loaded by the reference (and by sdslgpu_load_sdsl).  Same buffer protocol as
 * sdslgpu_bv_serialize.
 *   KIND_BV      what 0 / 1 / 2 = bit_vector / rank_support_v<1> / <0> (any handle; with SDSLGPU_F_SDSL_LAYOUT from the
 *                resident copies); what 5 / 6 = rank_support_v5<1> / <0> (rank_support_v5.hpp:66-158);
 *                what 3 / 4 = select_support_mcl<1> / <0>::serialize
 *                (select_support_mcl.hpp:474-518) with the contents of init_slow / init_fast (:207-381): the argument
 *                positions they store come from the batched select kernel, the host only packs them
 *   KIND_RRR63   what 0 = the complete rrr_vector<63>::serialize bytes (rrr_vector.hpp:366-378)
 *   KIND_SD      what 0 = size, wl, m_low, m_high (sd_vector.hpp:426-433; needs SDSLGPU_F_SDSL_LAYOUT);
 *                what 1 = the complete sd_vector<>::serialize bytes incl. the two select supports over m_high (:434-435)
 *   KIND_WT_HUFF what 0 = wt_pc::serialize (wt_pc.hpp:713-726): size, sigma, m_bv, rank_support_v<1>,
 *                select_support_mcl<1>, <0>, byte_tree (wt_helper.hpp:362-375); with SDSLGPU_F_RRR_BV the
 *                wt_huff<rrr_vector<63>> form (the rrr supports serialise to nothing, rrr_vector.hpp:580-585)
 *                what 1 (KIND_WT_HUFF and KIND_CSA_WT) = the same tree as wt_huff<bit_vector, rank_support_v5<>,
 *                select_support_scan<>, select_support_scan<0>>: with a CSA created at sa_dens = isa_dens = 2^20 this
 *                is the reference's count-benchmark index FM_HUFF (benchmark/indexing_count/index.config:8)
 *   KIND_WT_INT  what 0 = wt_int::serialize (wt_int.hpp:792-805)
 *   KIND_CSA_WT  what 0 = csa_wt::serialize (csa_wt.hpp:389-402): wavelet tree, SA samples, ISA samples (both
 *                int_vector<0> of width hi(size)+1, csa_sampling_strategy.hpp:103,762), byte_alphabet
 *                (csa_alphabet_strategy.hpp:258-268)
 * Byte-identical to the reference for every non-empty input (tests/test_egress_gpu.py; an EMPTY wt_huff of the
 * reference serialises uninitialised tables, here they are written as "no symbol"). */
int sdslgpu_serialize(const sdslgpu_handle *h, int what, void *buf, uint64_t cap, uint64_t *nbytes);

/* ---- multi-GPU groups ------------------------------------------------------------------------- */

/* One box, several B200s: the index is REPLICATED on every member of a group, a query batch is SHARDED over the
 * members (member r answers queries [r*s, (r+1)*s), s = n / nranks; the n - s*nranks left-over queries are answered by
 * everybody) and the results are ALL-GATHERED, so that after the call every member's `out` holds all n answers
 * (BASELINE.json north_star; SURVEY.md §8(b) lines 486-489, §8(e)).  The reference has no counterpart: its queries are
 * scalar const member calls (rank_support_v.hpp:129-139) that a host program spreads over threads itself.
 *
 * A group is created either
 *   - in ONE process driving several devices: sdslgpu_group_create(devices, ndev) -> a group with ndev LOCAL members
 *     (ncclCommInitAll underneath); or
 *   - with one process per GPU (torchrun, MPI): rank 0 calls sdslgpu_group_unique_id, ships the 128 bytes to the other
 *     ranks by any means, and every rank calls sdslgpu_group_create_rank (ncclCommInitRank) -> 1 local member each.
 * Every group call takes ARRAYS with one entry per LOCAL member (length ndev resp. 1): handles, device pointers,
 * streams.  Each member's idx array holds ALL n queries (the batch is identical on every member; only the member's
 * shard is read), each member's out array has room for n results.  All pointers are DEVICE pointers on the member's
 * device.  streams == NULL: the group's own streams are used and the call returns when the results are complete;
 * otherwise the call is asynchronous on streams[k] (an entry that is NULL means the legacy default stream).
 * Group calls are collective: every member (every rank) must make the same calls in the same order.
 * NCCL is bound at run time (dlopen of the libnccl.so.2 already in the process, else from the loader path or
 * $SDSLGPU_NCCL_LIB); a box without it gets SDSLGPU_ENOTSUP from the create calls.
 * Devices listed twice in sdslgpu_group_create give a "loopback" group (several members on one GPU, no NCCL; only
 * SDSLGPU_GATHER_FUSED / _NONE) — it exists so that the fused path can be tested on a single GPU. */
typedef struct sdslgpu_group sdslgpu_group;
#define SDSLGPU_UNIQUE_ID_BYTES 128
#define SDSLGPU_MAX_GROUP 16

#define SDSLGPU_GATHER_NONE 0  /* no gather: member r's out holds only its shard [r*s, (r+1)*s) and the left-over tail */
#define SDSLGPU_GATHER_NCCL 1  /* kernels, then ncclAllGather in place on out */
#define SDSLGPU_GATHER_FUSED 2 /* the shard's last kernel stores every result into ALL members' out arrays over NVLink
                                  (peer memory); out must come from sdslgpu_group_alloc.  Plain bit vectors: the un-sort
                                  stage of the binned pipeline does the stores; other ops: a peer-store copy kernel
                                  behind theirs.  No NCCL call on the data path. */
#define SDSLGPU_GATHER_AUTO 3  /* the fastest available: FUSED when out is group-allocated, else PACKED when peers can map
                                  each other's memory, else NCCL */
#define SDSLGPU_GATHER_PACKED 4 /* like FUSED, but the answers cross NVLink as w-bit fields (w = bits of the largest
                                  possible answer, e.g. 34 instead of 64 for a 2^33-bit vector: the int_vector<w>
                                  layout) into a staging buffer the group owns on every member, and a kernel on the
                                  receiving side widens them into out — which may be ANY device memory here.  Half the
                                  NVLink bytes of FUSED, but measured 10-17 % slower than FUSED (the 136-byte stores
                                  per warp and peer are not line-aligned, and the widening is one more kernel) and
                                  3-17 % faster than NCCL: the mode for result arrays the group did not allocate. */

int sdslgpu_group_unique_id(void *id128);
int sdslgpu_group_create_rank(const void *id128, int nranks, int rank, int device, sdslgpu_group **out);
int sdslgpu_group_create(const int *devices, int ndev, sdslgpu_group **out);
int sdslgpu_group_free(sdslgpu_group *g);
/* any of the out pointers may be NULL.  *fused_possible = 1 when the members can store into each other's memory */
int sdslgpu_group_info(const sdslgpu_group *g, int *nranks, int *nlocal, int *first_rank, int *fused_possible);

/* Symmetric device memory: `bytes` on every member (zero-filled), ptrs[k] = local member k's copy.  Memory from this
 * call is what SDSLGPU_GATHER_FUSED needs for `out` (every member must pass the same offset into it).  Collective. */
int sdslgpu_group_alloc(sdslgpu_group *g, uint64_t bytes, void **ptrs);
int sdslgpu_group_release(sdslgpu_group *g, void *const *ptrs);

/* Replicates the index behind `src` (given on the rank that owns global rank `root`, NULL elsewhere) onto every member:
 * out[k] = a new handle on local member k's device (the root's member gets its own copy too); free each with
 * sdslgpu_free.  The index travels in the reference's own serialised form (sdslgpu_serialize -> ncclBroadcast ->
 * sdslgpu_load_sdsl_ex), so a replica answers exactly like its source.  Collective. */
int sdslgpu_group_replicate(sdslgpu_group *g, const sdslgpu_handle *src, int root, sdslgpu_handle **out);

/* Sharded forms of sdslgpu_rank / sdslgpu_select (KIND_BV, KIND_RRR63, KIND_SD handles, b = 0 / 1; plain bit vectors
 * have the peer stores fused into their kernels, the compressed ones use the copy kernel), sdslgpu_wt_rank (byte
 * trees: KIND_WT_HUFF, KIND_CSA_WT) and sdslgpu_fm_count: same results, in the same order, on every member. */
int sdslgpu_group_rank(sdslgpu_group *g, const sdslgpu_handle *const *h, int b, const uint64_t *const *idx, uint64_t n,
                       uint64_t *const *out, int gather, void *const *streams);
int sdslgpu_group_select(sdslgpu_group *g, const sdslgpu_handle *const *h, int b, const uint64_t *const *i, uint64_t n,
                         uint64_t *const *out, int gather, void *const *streams);
int sdslgpu_group_wt_rank(sdslgpu_group *g, const sdslgpu_handle *const *h, const uint64_t *const *i, const uint8_t *const *c,
                          uint64_t n, uint64_t *const *out, int gather, void *const *streams);
int sdslgpu_group_fm_count(sdslgpu_group *g, const sdslgpu_handle *const *h, const uint8_t *const *pats,
                           const uint64_t *const *pat_off, uint64_t n, uint64_t *const *cnt_out, int gather,
                           void *const *streams);

#ifdef __cplusplus
}
#endif
#endif /* SDSLGPU_H */
