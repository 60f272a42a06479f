// internal.h — host-side handle layout and helpers shared by the C-ABI translation units.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/sdslgpu.h"
#include "bits_access.cuh"
#include "bv_device.cuh"
#include "common.cuh"
#include "wt_tree.h"

namespace sdslgpu
{

void set_error(char const * fmt, ...);
int cuda_fail(cudaError_t e, char const * what, char const * file, int line);

#define SG_CUDA(expr)                                                                                                  \
    do                                                                                                                 \
    {                                                                                                                  \
        cudaError_t e__ = (expr);                                                                                      \
        if (e__ != cudaSuccess)                                                                                        \
            return ::sdslgpu::cuda_fail(e__, #expr, __FILE__, __LINE__);                                               \
    } while (0)

#define SG_TRY(expr)                                                                                                   \
    do                                                                                                                 \
    {                                                                                                                  \
        int s__ = (expr);                                                                                              \
        if (s__ != SDSLGPU_OK)                                                                                         \
            return s__;                                                                                                \
    } while (0)

// grid sizing: query kernels are latency-bound gathers, so run a grid-stride loop over sm_count() x 8 resident
// CTAs of 256 threads (full occupancy at <= 32 registers/thread).  sm_count() = cudaDevAttrMultiProcessorCount of the
// CURRENT device (148 on a B200), queried once per device and cached (api.cu).
int sm_count();
static constexpr int kThreads = 256;

struct DeviceGuard
{
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev)
    {
        if (cudaGetDevice(&prev) == cudaSuccess && cudaSetDevice(dev) == cudaSuccess)
            ok = true;
    }
    ~DeviceGuard()
    {
        if (prev >= 0)
            cudaSetDevice(prev);
    }
};

// Owns device allocations of one handle and keeps the byte total.
struct DevicePool
{
    std::vector<void *> ptrs;
    std::vector<uint64_t> sizes_;
    uint64_t bytes = 0;
    int alloc(void ** p, uint64_t n);
    template <class T>
    int alloc_t(T ** p, uint64_t count)
    {
        return alloc(reinterpret_cast<void **>(p), count * sizeof(T));
    }
    void release(void * p); // free one allocation early (build temporaries)
    void release_all();
};

// Pinned-free staging path for HOST query buffers: device chunk buffers + two streams, so that the
// H2D copy of chunk k+1, the kernel of chunk k and the D2H copy of chunk k-1 overlap (PCIe is full duplex).
struct Staging
{
    static constexpr int kSlots = 3;
    static constexpr uint64_t kChunk = 1ull << 22; // queries per chunk
    std::mutex mu;
    bool ready = false;
    uint8_t * in[kSlots] = {nullptr, nullptr, nullptr};  // kChunk * in_bytes_max
    uint8_t * out[kSlots] = {nullptr, nullptr, nullptr}; // kChunk * out_bytes_max
    cudaStream_t stream[kSlots] = {nullptr, nullptr, nullptr};
    static constexpr uint64_t kInBytesPerQuery = 16; // largest input record (i:u64 + c:u64)
    static constexpr uint64_t kOutBytesPerQuery = 16;
    int ensure();
    void destroy();
};

// Result fan-out of the multi-GPU group calls (group.cu): a kernel that writes out[p] also stores the same value to
// dst[r][p] for r < n — the `out` arrays of the other members of the group, mapped through CUDA IPC / peer access, so
// the all-gather of results rides on the kernel's own stores over NVLink instead of a separate collective.
static constexpr int kMaxFan = 15;
struct Fan
{
    uint64_t * dst[kMaxFan];
    uint32_t n = 0;
    uint32_t width = 0; // 0: dst[r] are the peers' u64 result arrays; w: their staging regions for w-bit fields (fan.cuh)
};

// Staging of the int_vector<w> wire format (sdslgpu_rank_iv / _select_iv): per slot the packed chunk as it crosses
// PCIe, its unpacked u64 form for the kernels, and the same two for the results.
struct StagingIv
{
    static constexpr int kSlots = 3;
    static constexpr uint64_t kChunk = 1ull << 23; // queries per chunk (a multiple of 64: chunks start on word boundaries)
    bool ready = false;
    uint64_t * pin[kSlots] = {nullptr, nullptr, nullptr};  // packed queries  (kChunk words: width <= 64)
    uint64_t * pout[kSlots] = {nullptr, nullptr, nullptr}; // packed results
    uint64_t * uin[kSlots] = {nullptr, nullptr, nullptr};  // unpacked queries
    uint64_t * uout[kSlots] = {nullptr, nullptr, nullptr}; // unpacked results
    cudaStream_t stream[kSlots] = {nullptr, nullptr, nullptr};
    int ensure();
    void destroy();
};

enum class PtrSpace
{
    Host,
    Device
};
// classifies a user pointer; device pointers must belong to `device`
int classify(void const * p, int device, PtrSpace * space);

} // namespace sdslgpu

// ------------------------------------------------------------------------------------------------
// Device images.  A BvImage is the building block: wavelet trees and the Elias-Fano high part own one.
// ------------------------------------------------------------------------------------------------
namespace sdslgpu
{

struct BvImage
{
    uint64_t nbits = 0;
    uint64_t nblocks = 0;      // nbits/224 + 1 (one past the end so rank(size) stays in range)
    bvblock * blocks = nullptr; // sector-interleaved payload + counts
    uint64_t * top = nullptr;   // absolute 1-count per superblock of 2^24 blocks
    uint64_t ntop = 0;
    uint64_t ones = 0;
    int order = SDSLGPU_ORDER_AUTO; // sdslgpu_set_batch_order: direct / binned execution of large batches (binned.cu)
    // select samples, per pattern b: samp[b][j] = block holding the (j*S+1)-th b-bit; two sentinels
    uint32_t * samp[2] = {nullptr, nullptr};
    uint64_t nsamp[2] = {0, 0};
    uint32_t log_s[2] = {6, 6};
    uint32_t interp[2] = {0, 0}; // interpolate between samples (set when the stride exceeds 64)
    uint32_t samp_pos[2] = {0, 0}; // samples hold (position >> 5) instead of block indices (vectors up to 2^36 bits)
    // select sectors (bv_device.cuh): built by bv_ensure_select_sectors on the first large select batch of a KIND_BV handle
    bvblock * sect[2] = {nullptr, nullptr};
    uint64_t nsect[2] = {0, 0};
    uint32_t sect_stride[2] = {0, 0};
    uint64_t sect_magic[2] = {0, 0};
    bool sect_tried[2] = {false, false}; // built, or found not to apply (density, size, memory): do not try again
    // optional SDSL layout (SDSLGPU_F_SDSL_LAYOUT): raw words (+ pad) and the m_basic_block tables
    uint64_t * words = nullptr;
    uint64_t nwords = 0;
    uint64_t * rank_table[2] = {nullptr, nullptr}; // [b]
    uint64_t table_words = 0;
};

inline BvView bv_view(BvImage const & v)
{
    BvView w;
    w.blocks = v.blocks;
    w.top = v.top;
    w.samp[0] = v.samp[0];
    w.samp[1] = v.samp[1];
    w.log_s[0] = v.log_s[0];
    w.log_s[1] = v.log_s[1];
    w.interp[0] = v.interp[0];
    w.interp[1] = v.interp[1];
    w.samp_pos[0] = v.samp_pos[0];
    w.samp_pos[1] = v.samp_pos[1];
    w.sect[0] = v.sect[0];
    w.sect[1] = v.sect[1];
    w.sect_stride[0] = v.sect_stride[0];
    w.sect_stride[1] = v.sect_stride[1];
    w.sect_magic[0] = v.sect_magic[0];
    w.sect_magic[1] = v.sect_magic[1];
    w.nbits = v.nbits;
    w.ones = v.ones;
    return w;
}

// rrr_vector<63, int_vector<>, 32> (rrr_vector.hpp:101-109)
struct RrrImage
{
    uint64_t size = 0, nblocks = 0, nsuper = 0, ones = 0, btnr_bits = 0;
    uint64_t * btnr = nullptr;    // m_btnr
    uint64_t * records = nullptr; // 64-byte record per superblock (rank, btnrp|invert, 32 classes, quarter sums) + closing record
    void * tables = nullptr;      // RrrTables (binomials + code lengths), device copy
    uint32_t * hint[2] = {nullptr, nullptr}; // select hints: superblock of every 2^hint_shift[b]-th b-bit
    uint32_t hint_shift[2] = {13, 13};
};

struct WtHuffImage
{
    uint64_t size = 0, sigma = 0;
    BvImage bv;               // the single concatenated bit vector m_bv (wt_pc.hpp:88-94) as sector blocks ...
    RrrImage rrr;             // ... or H0-compressed (wt_huff<rrr_vector<63>>) when use_rrr
    bool use_rrr = false;
    WtTree * tree = nullptr;   // device copy
    WtTree host_tree;          // host copy (kept for serialisation / introspection)
};

// wt_int<> (wt_int.hpp:85-91): max_level levels of `size` bits each, concatenated
struct WtIntImage
{
    uint64_t size = 0, sigma = 0;
    uint32_t max_level = 0;
    BvImage tree;
};

// byte_alphabet of a CSA (csa_alphabet_strategy.hpp:136-212), staged in shared memory next to the tree
struct alignas(16) FmTables
{
    uint64_t C[257];
    uint8_t char2comp[256];
    uint8_t comp2char[256];
    uint32_t sigma;
    uint32_t pad_;
};

// sd_vector<> (sd_vector.hpp:155-163)
struct SdImage
{
    uint64_t size = 0, m = 0, high_bits = 0, low_words = 0;
    uint32_t wl = 0;
    uint64_t * low = nullptr; // m_low, packed wl-bit entries
    BvImage high;             // m_high with rank blocks + select<1>/<0> samples
    // select_0 of the vector itself (the job of select_0_support_sd, sd_vector.hpp:752-921): block of `high` in which
    // every 2^log_s0-th zero of the vector is crossed (sd_device.cuh sd_select0_one)
    uint32_t * samp0 = nullptr;
    uint64_t nsamp0 = 0;
    uint32_t log_s0 = 0;
};

// one-hot occurrence bitmaps of the BWT (occ16_device.cuh), staged in shared memory by the fm16 kernels
struct alignas(16) Occ16Tab
{
    bvblock const * blocks[2][16];
    uint64_t const * top[2][16];
    uint64_t CH[16];       // level-1 start of the subsequence of each high nibble
    uint64_t D[256];       // per comp symbol: C[cc] - rank1(B1[lo], CH[h])  (wrapping)
    uint8_t const * bwtc;  // the BWT in comp codes
    uint32_t levels;       // 1 (sigma <= 16) or 2
    uint32_t pad_;
};
static_assert(sizeof(Occ16Tab) % 16 == 0, "staged with 16-byte copies");

struct Occ16Image
{
    uint32_t levels = 0; // 0: not built (SDSLGPU_F_COMPACT / rrr-backed / ingest without it)
    BvImage bm[2][16];
    uint8_t * bwtc = nullptr;
    Occ16Tab * tab = nullptr; // device copy
    Occ16Tab host_tab;
};

struct CsaImage
{
    Occ16Image occ;
    uint64_t n = 0;              // csa.size() = text length + 1
    uint32_t sa_dens = 32;       // t_dens (csa_wt.hpp:50)
    uint64_t * samples = nullptr; // SA[0], SA[dens], ... widened to u64 (csa_sampling_strategy.hpp:98-115)
    uint64_t nsamples = 0;
    uint32_t isa_dens = 64;           // t_inv_dens (csa_wt.hpp:51)
    uint64_t * isa_samples = nullptr; // ISA[0], ISA[64], ... widened to u64 (csa_sampling_strategy.hpp:758-779)
    uint64_t nisa = 0;
    FmTables * tab = nullptr; // device copy
    FmTables host_tab;
};

} // namespace sdslgpu

struct sdslgpu_handle
{
    int kind = 0;
    int device = 0;
    uint32_t flags = 0;
    int order = SDSLGPU_ORDER_AUTO; // sdslgpu_set_batch_order (bit vectors keep their own copy in BvImage::order)
    sdslgpu::DevicePool pool;
    sdslgpu::Staging staging;
    sdslgpu::StagingIv staging_iv;
    sdslgpu::BvImage bv;        // KIND_BV
    // KIND_BV: indicator vectors of the two-bit patterns 10 / 01 / 00 / 11, built on first use (bv.cu)
    sdslgpu::BvImage pat[4];
    bool pat_ready[4] = {false, false, false, false};
    std::mutex pat_mu;
    sdslgpu::WtHuffImage wt;    // KIND_WT_HUFF (and the BWT of KIND_CSA_WT)
    sdslgpu::CsaImage csa;      // KIND_CSA_WT
    sdslgpu::WtIntImage wti;    // KIND_WT_INT
    sdslgpu::RrrImage rrr;      // KIND_RRR63
    sdslgpu::SdImage sd;        // KIND_SD
    // sdslgpu_serialize: the blob computed by a size query (buf == NULL) is kept for the call that fetches it
    std::mutex ser_mu;
    std::vector<uint8_t> ser_blob;
    int ser_what = -1;
};

namespace sdslgpu
{
inline RrrView rrr_view(RrrImage const & r)
{
    RrrView v;
    v.size = r.size;
    v.nblocks = r.nblocks;
    v.nsuper = r.nsuper;
    v.ones = r.ones;
    v.btnr = r.btnr;
    v.records = r.records;
    v.tables = reinterpret_cast<RrrTables const *>(r.tables);
    v.hint[0] = r.hint[0];
    v.hint[1] = r.hint[1];
    v.hint_shift[0] = r.hint_shift[0];
    v.hint_shift[1] = r.hint_shift[1];
    v.try_sparse = (r.ones * 32 <= r.size || (r.size - r.ones) * 32 <= r.size) ? 1u : 0u;
    return v;
}
inline PlainBits plain_bits(WtHuffImage const & w)
{
    PlainBits b;
    b.v = bv_view(w.bv);
    return b;
}
inline RrrBits rrr_bits(WtHuffImage const & w)
{
    RrrBits b;
    b.v = rrr_view(w.rrr);
    b.t = nullptr;
    return b;
}
// launches KERNEL<PlainBits> or KERNEL<RrrBits> (with the binomial tables appended to the dynamic shared memory)
#define SG_LAUNCH_BITS(KERNEL, W, GRID, BASE_SMEM, STREAM, ...)                                                        \
    do                                                                                                                 \
    {                                                                                                                  \
        if ((W).use_rrr)                                                                                               \
        {                                                                                                              \
            auto kfn__ = KERNEL<RrrBits>;                                                                              \
            cudaFuncSetAttribute(kfn__, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((BASE_SMEM) + RrrBits::kSmem)); \
            kfn__<<<(GRID), kThreads, (BASE_SMEM) + RrrBits::kSmem, (STREAM)>>>(rrr_bits(W), __VA_ARGS__);             \
        }                                                                                                              \
        else                                                                                                           \
            KERNEL<PlainBits><<<(GRID), kThreads, (BASE_SMEM), (STREAM)>>>(plain_bits(W), __VA_ARGS__);                \
    } while (0)

// bv.cu
int bv_build(DevicePool & pool, BvImage & v, uint32_t flags, uint64_t const * words_host_or_dev, bool words_on_device, uint64_t nbits, cudaStream_t s);
int bv_build_sdsl_rank_table(DevicePool & pool, BvImage & v, int b, cudaStream_t s, bool v5 = false);
int bv_build_pattern(DevicePool & pool, BvImage const & src, int pat, BvImage & dst, cudaStream_t s);
// fan / fanned: optional result fan-out (see Fan); *fanned tells whether the kernels that ran did the remote stores
// (the binned pipeline's un-sort does; after a direct kernel the caller copies the shard itself)
int bv_rank_device(BvImage const & v, uint32_t flags, int b, uint64_t const * idx, uint64_t n, uint64_t * out, cudaStream_t s, Fan const * fan = nullptr,
                   bool * fanned = nullptr);
int bv_select_device(BvImage const & v, int b, uint64_t const * idx, uint64_t n, uint64_t * out, cudaStream_t s, Fan const * fan = nullptr,
                     bool * fanned = nullptr);
int bv_access_device(BvImage const & v, uint64_t const * idx, uint64_t n, uint64_t * out, cudaStream_t s);
// binned.cu: locality-ordered execution of large batches; *done = false means "not applicable, use the direct kernel"
// bytes of index per query at the break-even between the two batch orders (binned.cu bin_wanted): one-gather ops (rank,
// select through select sectors) / ops with two or more dependent gathers (sampled select, sd and rrr selects)
static constexpr uint32_t kBinRankDensity = 128, kBinSelectDensity = 192;
bool bv_binned_wanted(BvImage const & v, uint64_t n, bool select = false, int b = 1);
bool bin_wanted(int order, uint64_t index_bytes, uint64_t n, uint32_t index_bytes_per_query = kBinRankDensity);
int bv_rank_binned_device(BvImage const & v, int b, uint64_t const * idx, uint64_t n, uint64_t * out, cudaStream_t s, bool * done, Fan const * fan = nullptr);
int bv_ensure_select_sectors(sdslgpu_handle const * h, int b, uint64_t n); // bv.cu; a no-op unless a binned select of n queries would use them
int bv_ensure_select_sectors_image(sdslgpu_handle const * h, BvImage const & v, int b, uint64_t reserve_bytes); // any bit-vector image the handle owns
int bv_select_binned_device(BvImage const & v, int b, uint64_t const * idx, uint64_t n, uint64_t * out, cudaStream_t s, bool * done, Fan const * fan = nullptr);
// wt.cu
int wt_huff_upload(sdslgpu_handle * h, uint64_t size, uint64_t sigma, WtTree const & tree, uint64_t const * bv_words, uint64_t bv_bits, cudaStream_t s);
int wt_huff_finish(sdslgpu_handle * h, uint64_t size, uint64_t sigma, WtTree const & tree, cudaStream_t s);
int wt_huff_build_from_text(sdslgpu_handle * h, uint8_t const * text_host, uint64_t n, cudaStream_t s);
int wt_rank_device(sdslgpu_handle const * h, uint64_t const * i, uint8_t const * c, uint64_t n, uint64_t * out, cudaStream_t s);
int wt_select_device(sdslgpu_handle const * h, uint64_t const * i, uint8_t const * c, uint64_t n, uint64_t * out, cudaStream_t s);
int wt_access_device(sdslgpu_handle const * h, uint64_t const * i, uint64_t n, uint64_t * sym, uint64_t * rnk, cudaStream_t s);
// wt_build.cu: bit planes on the device; SDSLGPU_ENOTSUP = not enough device memory for the scratch (use the host fill)
int wt_histogram_device(uint8_t const * d_text, uint64_t n, uint64_t (&C)[256], cudaStream_t s);
int wt_huff_planes_device(uint8_t const * d_text, uint64_t n, WtTree const & tree, uint64_t bits, uint64_t * d_words, cudaStream_t s);
int wt_int_planes_device(uint64_t * d_seq, uint64_t n, uint32_t * max_level_out, uint64_t * sigma_out, uint64_t ** d_words_out, cudaStream_t s);
// wt_int.cu
int wt_int_build(sdslgpu_handle * h, uint64_t const * seq_host, uint64_t n, cudaStream_t s);
int wt_int_rank_device(sdslgpu_handle const * h, uint64_t const * i, uint64_t const * c, uint64_t n, uint64_t * out, cudaStream_t s);
int wt_int_select_device(sdslgpu_handle const * h, uint64_t const * i, uint64_t const * c, uint64_t n, uint64_t * out, cudaStream_t s);
int wt_int_access_device(sdslgpu_handle const * h, uint64_t const * i, uint64_t n, uint64_t * sym, uint64_t * rnk, cudaStream_t s);
// rrr.cu
int rrr_build(sdslgpu_handle * h, uint64_t const * words_host_or_dev, bool on_device, uint64_t nbits, cudaStream_t s);
int rrr_rank_device(sdslgpu_handle const * h, int b, uint64_t const * idx, uint64_t n, uint64_t * out, cudaStream_t s);
int rrr_select_device(sdslgpu_handle const * h, int b, uint64_t const * idx, uint64_t n, uint64_t * out, cudaStream_t s);
int rrr_access_device(sdslgpu_handle const * h, uint64_t const * idx, uint64_t n, uint64_t * out, cudaStream_t s);
int rrr_serialize(sdslgpu_handle const * h, std::vector<uint8_t> & blob);
int rrr_serialize_image(RrrImage const & r, std::vector<uint8_t> & blob); // appends rrr_vector<63>::serialize bytes
int rrr_upload_tables(DevicePool & pool, RrrImage & r, cudaStream_t s);
int rrr_build_hints(DevicePool & pool, RrrImage & r, cudaStream_t s);
int rrr_build_image(DevicePool & pool, RrrImage & r, uint64_t const * words_host_or_dev, bool on_device, uint64_t nbits, cudaStream_t s);
int rrr_rank_image(RrrImage const & r, int b, uint64_t const * idx, uint64_t n, uint64_t * out, cudaStream_t s, int order = SDSLGPU_ORDER_DIRECT);
int rrr_records_from_sdsl(DevicePool & pool, RrrImage & r, uint64_t const * bt_words, uint64_t nblocks, std::vector<uint64_t> const & rank,
                          std::vector<uint64_t> const & btnrp, std::vector<uint8_t> const & invert, uint64_t total_bits_hint, cudaStream_t s);
// sdsl_format.cu
int load_sdsl_blob(sdslgpu_handle * h, uint8_t const * blob, uint64_t nbytes, uint32_t sa_dens, uint32_t isa_dens, uint64_t * consumed, cudaStream_t s);
// sd.cu
int sd_build_select0_samples(sdslgpu_handle * h, cudaStream_t s);
int sd_build(sdslgpu_handle * h, uint64_t const * words_host_or_dev, bool on_device, uint64_t nbits, cudaStream_t s);
int sd_rank_device(sdslgpu_handle const * h, int b, uint64_t const * idx, uint64_t n, uint64_t * out, cudaStream_t s);
int sd_select_device(sdslgpu_handle const * h, int b, uint64_t const * idx, uint64_t n, uint64_t * out, cudaStream_t s);
int sd_access_device(sdslgpu_handle const * h, uint64_t const * idx, uint64_t n, uint64_t * out, cudaStream_t s);
int sd_serialize_low_high(sdslgpu_handle const * h, std::vector<uint8_t> & blob);
// sdsl_egress.cu: complete reference-format blobs (what the reference's serialize() / store_to_file writes)
int egress_bv_part(BvImage const & v, int what, std::vector<uint8_t> & blob);
int egress_select_mcl(BvImage const & v, int b, std::vector<uint8_t> & blob);
int egress_sd(sdslgpu_handle const * h, std::vector<uint8_t> & blob);
int egress_wt_huff(sdslgpu_handle const * h, std::vector<uint8_t> & blob, bool v5_scan = false);
int egress_wt_int(sdslgpu_handle const * h, std::vector<uint8_t> & blob);
int egress_csa(sdslgpu_handle const * h, std::vector<uint8_t> & blob, bool v5_scan = false);
// gpu_sa.cu
int gpu_suffix_array_bwt(uint8_t const * text_host, uint64_t len, uint32_t dens, uint32_t isa_dens, std::vector<uint8_t> & bwt, std::vector<uint64_t> & samples,
                         std::vector<uint64_t> & isa_samples, uint32_t * rounds_out, cudaStream_t s);
// fm.cu
int csa_build_from_text(sdslgpu_handle * h, uint8_t const * text_host, uint64_t len, uint32_t sa_dens, uint32_t isa_dens, cudaStream_t s);
int csa_upload(sdslgpu_handle * h, uint8_t const * bwt_host, uint64_t const * samples_host, uint64_t nsamples, uint64_t const * isa_host, uint64_t nisa, cudaStream_t s);
int csa_upload_isa(sdslgpu_handle * h, uint64_t const * isa_host, uint64_t nisa, cudaStream_t s);
int fm_extract_device(sdslgpu_handle const * h, uint64_t const * begin, uint64_t const * end, uint64_t const * out_off, uint64_t n, uint8_t * out, cudaStream_t s);
int fm_count_device(sdslgpu_handle const * h, uint8_t const * pats, uint64_t const * off, uint64_t npat, uint64_t * cnt, uint64_t * l, cudaStream_t s);
int fm_sa_device(sdslgpu_handle const * h, uint64_t const * idx, uint64_t cnt, uint64_t * out, cudaStream_t s);
int fm_scan_counts_device(uint64_t const * cnt, uint64_t npat, uint64_t * occ_off, uint64_t * tmp, cudaStream_t s);
int fm_locate_fill_device(sdslgpu_handle const * h, uint64_t const * l, uint64_t const * occ_off, uint64_t npat, uint64_t total, uint64_t * occ, cudaStream_t s);
uint64_t fm_scan_tmp_words(uint64_t npat);
// fm16.cu: the same four searches over the one-hot occurrence structure
int occ16_build(sdslgpu_handle * h, uint8_t const * bwt_host, cudaStream_t s);
int occ16_build_from_wt(sdslgpu_handle * h, cudaStream_t s);
int fm16_count_device(sdslgpu_handle const * h, uint8_t const * pats, uint64_t const * off, uint64_t npat, uint64_t * cnt, uint64_t * l, cudaStream_t s);
int fm16_sa_device(sdslgpu_handle const * h, uint64_t const * idx, uint64_t cnt, uint64_t * out, cudaStream_t s);
int fm16_locate_fill_device(sdslgpu_handle const * h, uint64_t const * l, uint64_t const * occ_off, uint64_t npat, uint64_t total, uint64_t * occ, cudaStream_t s);
int fm16_extract_device(sdslgpu_handle const * h, uint64_t const * begin, uint64_t const * end, uint64_t const * out_off, uint64_t n, uint8_t * out, cudaStream_t s);
unsigned grid_for(uint64_t n, int per_thread = 1);
unsigned blocks_for(uint64_t n);
} // namespace sdslgpu

static_assert(sizeof(sdslgpu::WtTree) % 16 == 0 && sizeof(sdslgpu::FmTables) % 16 == 0, "tables are staged with 16-byte copies");
