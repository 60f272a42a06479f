// rrr.cu — rrr_vector<63, int_vector<>, 32>: device-side encoder and batched rank / select / access.
//
// Replaces (results bit-exact; m_bt and m_btnr are bit-identical to the reference's, so SDSL-serialised
// vectors can be ingested as they are):
//   rrr_vector ctor            rrr_vector.hpp:158-270 + rrr_helper.hpp:346-366 (bin_to_nr) -> rrr_classify /
//                                                                  rrr_superblock / rrr_encode kernels + scans
//   rank_support_rrr::rank     rrr_vector.hpp:503-544  -> rrr_rank_kernel
//   select_support_rrr::select rrr_vector.hpp:639-726  -> rrr_select_kernel<B>
//   rrr_vector::operator[]     rrr_vector.hpp:276-298  -> rrr_access_kernel
// Device layout: the 6-bit classes (m_bt) and the offsets (m_btnr) stay packed exactly as in the reference;
// the three per-superblock arrays (m_rank, m_btnrp, m_invert) are fused into one 16-byte record so a query
// reads them with a single 128-bit load.  C(n,k) for n,k <= 63 (32 KB) and the code lengths live in shared memory.
#include "internal.h"
#include "scan.cuh"

namespace sdslgpu
{

static constexpr uint32_t kBs = 63; // t_bs
static constexpr uint32_t kK = 32;  // t_k
static constexpr uint64_t kInvBit = 1ull << 63;

struct RrrTables
{
    uint64_t binom[64][64]; // binom[n][k] = C(n, k), 0 for k > n   (rrr_helper.hpp:193-237)
    uint8_t space[64];      // bits of an offset of class k: 0 if C(63,k) == 1 else hi(C(63,k)) + 1 (:286-293)
};

static RrrTables const & host_tables()
{
    static RrrTables t;
    static bool ready = false;
    if (!ready)
    {
        std::memset(&t, 0, sizeof(t));
        uint64_t full[65][65];
        std::memset(full, 0, sizeof(full));
        for (int n = 0; n <= 64; ++n)
            full[n][0] = 1;
        for (int n = 1; n <= 64; ++n)
            for (int k = 1; k <= n; ++k)
                full[n][k] = full[n - 1][k - 1] + full[n - 1][k];
        for (int n = 0; n < 64; ++n)
            for (int k = 0; k < 64; ++k)
                t.binom[n][k] = full[n][k];
        for (int k = 0; k < 64; ++k)
        {
            uint64_t c = full[63][k];
            uint8_t hi = 0;
            for (uint64_t x = c; x >>= 1;)
                ++hi;
            t.space[k] = (c == 1) ? 0 : (uint8_t)(hi + 1);
        }
        ready = true;
    }
    return t;
}

__device__ __forceinline__ void stage_rrr(RrrTables const * __restrict__ g, RrrTables * s)
{
    uint4 const * src = reinterpret_cast<uint4 const *>(g);
    uint4 * dst = reinterpret_cast<uint4 *>(s);
    for (uint32_t k = threadIdx.x; k < sizeof(RrrTables) / 16; k += blockDim.x)
        dst[k] = __ldg(src + k);
    __syncthreads();
}

// the block with k ones and offset nr, decoded up to `upto` positions (inverse of bin_to_nr, rrr_helper.hpp:346-366;
// what decode_bit / decode_popcount / decode_select of rrr_helper.hpp:369-649 all compute from)
__device__ __forceinline__ uint64_t rrr_decode(RrrTables const * t, uint32_t k, uint64_t nr, uint32_t upto)
{
    if (k == 0)
        return 0;
    if (k == kBs)
        return (1ull << kBs) - 1;
    uint64_t bin = 0;
    for (uint32_t p = 0; p < upto && k; ++p)
    {
        uint64_t c = t->binom[kBs - 1 - p][k];
        if (nr >= c)
        {
            nr -= c;
            bin |= 1ull << p;
            --k;
        }
    }
    return bin;
}

struct RrrView
{
    uint64_t size;
    uint64_t nblocks;  // m_bt.size()
    uint64_t nsuper;   // m_btnrp.size(); records has nsuper + 1 entries (the last holds the total)
    uint64_t ones;
    uint64_t const * bt;      // packed 6-bit stored classes
    uint64_t const * btnr;    // packed offsets
    uint64_t const * records; // 2 words per superblock: ones before it, (bit offset into btnr) | invert << 63
    RrrTables const * tables;
    uint32_t const * hint[2]; // hint[b][j] = superblock holding the (j * 2^kHintShift + 1)-th b-bit (+ sentinels)
};

static constexpr uint32_t kHintShift = 13;

__device__ __forceinline__ uint32_t rrr_class(uint64_t const * __restrict__ bt, uint64_t j)
{
    return (uint32_t)read_int(bt, j * 6, 6);
}

// classes of superblock g up to (not including) block `blk`: ones and btnr bits consumed
__device__ __forceinline__ void rrr_scan_classes(RrrView const & v, RrrTables const * t, uint64_t g, uint32_t nblk, bool inv, uint64_t & ones, uint64_t & p)
{
    // 32 classes = 192 bits = exactly three words
    uint64_t const * w = v.bt + g * 3;
    uint64_t w0 = __ldg(w), w1 = nblk > 10 ? __ldg(w + 1) : 0, w2 = nblk > 21 ? __ldg(w + 2) : 0;
    for (uint32_t j = 0; j < nblk; ++j)
    {
        uint32_t bit = j * 6, c;
        if (bit < 60)
            c = (uint32_t)(w0 >> bit) & 63u;
        else if (bit == 60)
            c = (uint32_t)((w0 >> 60) | (w1 << 4)) & 63u;
        else if (bit < 124)
            c = (uint32_t)(w1 >> (bit - 64)) & 63u;
        else if (bit == 126)
            c = (uint32_t)((w1 >> 62) | (w2 << 2)) & 63u;
        else
            c = (uint32_t)(w2 >> (bit - 128)) & 63u;
        ones += inv ? kBs - c : c;
        p += t->space[c];
    }
}

// ------------------------------------------------------------------------------------------------
// queries
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t rrr_rank1_one(RrrView const & v, RrrTables const * t, uint64_t i)
{
    uint64_t blk = i / kBs, g = blk / kK;
    uint32_t off = (uint32_t)(i - blk * kBs);
    uint64_t r0, pw, r1;
    ld_pair(v.records + 2 * g, r0, pw);
    r1 = __ldg(v.records + 2 * g + 2);
    uint64_t d = r1 - r0;
    if (d == 0)
        return r0; // uniform superblocks (rrr_vector.hpp:514-523); same result as the general path
    if (d == (uint64_t)kBs * kK)
        return r0 + i - g * kK * kBs;
    bool inv = (pw & kInvBit) != 0;
    uint64_t p = pw & ~kInvBit, ones = r0;
    rrr_scan_classes(v, t, g, (uint32_t)(blk - g * kK), inv, ones, p);
    if (off == 0)
        return ones;
    uint32_t k = rrr_class(v.bt, blk);
    if (inv)
        k = kBs - k;
    uint32_t sp = t->space[k];
    uint64_t nr = sp ? read_int(v.btnr, p, sp) : 0;
    uint64_t bin = rrr_decode(t, k, nr, off);
    return ones + __popcll(bin & ((1ull << off) - 1));
}

__global__ void __launch_bounds__(kThreads) rrr_rank_kernel(RrrView const v, int b, uint64_t const * __restrict__ idx, uint64_t n, uint64_t * __restrict__ out)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RrrTables * t = reinterpret_cast<RrrTables *>(smem_raw);
    stage_rrr(v.tables, t);
    uint64_t const stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += stride)
    {
        uint64_t i = ld_stream_u64(idx + q);
        uint64_t r = SDSLGPU_NPOS;
        if (i <= v.size)
        {
            r = rrr_rank1_one(v, t, i);
            if (!b)
                r = i - r;
        }
        st_stream_u64(out + q, r);
    }
}

__global__ void __launch_bounds__(kThreads) rrr_access_kernel(RrrView const v, uint64_t const * __restrict__ idx, uint64_t n, uint64_t * __restrict__ out)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RrrTables * t = reinterpret_cast<RrrTables *>(smem_raw);
    stage_rrr(v.tables, t);
    uint64_t const stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += stride)
    {
        uint64_t i = ld_stream_u64(idx + q);
        uint64_t r = SDSLGPU_NPOS;
        if (i < v.size)
        {
            uint64_t blk = i / kBs, g = blk / kK;
            uint32_t off = (uint32_t)(i - blk * kBs);
            uint64_t r0, pw;
            ld_pair(v.records + 2 * g, r0, pw);
            bool inv = (pw & kInvBit) != 0;
            uint32_t k = rrr_class(v.bt, blk);
            if (inv)
                k = kBs - k;
            if (k == 0 || k == kBs)
                r = k != 0; // rrr_vector.hpp:283-288
            else
            {
                uint64_t p = pw & ~kInvBit, ones = 0;
                rrr_scan_classes(v, t, g, (uint32_t)(blk - g * kK), inv, ones, p);
                uint64_t bin = rrr_decode(t, k, read_int(v.btnr, p, t->space[k]), off + 1);
                r = (bin >> off) & 1;
            }
        }
        st_stream_u64(out + q, r);
    }
}

template <int B>
__global__ void __launch_bounds__(kThreads) rrr_select_kernel(RrrView const v, uint64_t const * __restrict__ idx, uint64_t n, uint64_t * __restrict__ out)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RrrTables * t = reinterpret_cast<RrrTables *>(smem_raw);
    stage_rrr(v.tables, t);
    uint64_t const args = B ? v.ones : v.size - v.ones;
    uint64_t const stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += stride)
    {
        uint64_t i = ld_stream_u64(idx + q);
        uint64_t r;
        if (i == 0)
            r = SDSLGPU_NPOS;
        else if (i > args)
            r = v.size; // the reference's in-band answer (rrr_vector.hpp:641-642, 686-689)
        else
        {
            // superblock g with count_before(g) < i <= count_before(g + 1)   (:643-655)
            uint64_t hj = (i - 1) >> kHintShift;
            uint64_t begin = __ldg(v.hint[B] + hj), end = (uint64_t)__ldg(v.hint[B] + hj + 1) + 1;
            while (end - begin > 1)
            {
                uint64_t mid = (begin + end) >> 1;
                uint64_t rk = __ldg(v.records + 2 * mid);
                uint64_t c = B ? rk : mid * kBs * kK - rk;
                if (c >= i)
                    end = mid;
                else
                    begin = mid;
            }
            uint64_t r0, pw;
            ld_pair(v.records + 2 * begin, r0, pw);
            uint64_t r1 = __ldg(v.records + 2 * begin + 2);
            uint64_t cnt = B ? r0 : begin * kBs * kK - r0;
            uint64_t d = r1 - r0;
            if (B ? (d == (uint64_t)kBs * kK) : (d == 0))
                r = begin * kK * kBs + (i - cnt - 1); // all-ones / all-zeros superblock (:658-663, :703-706)
            else
            {
                bool inv = (pw & kInvBit) != 0;
                uint64_t p = pw & ~kInvBit;
                uint64_t blk = begin * kK;
                uint32_t k = 0, sp = 0;
                for (;; ++blk)
                {
                    k = rrr_class(v.bt, blk);
                    if (inv)
                        k = kBs - k;
                    sp = t->space[k];
                    uint32_t c = B ? k : kBs - k;
                    if (cnt + c >= i)
                        break;
                    cnt += c;
                    p += sp;
                }
                uint64_t bin = rrr_decode(t, k, sp ? read_int(v.btnr, p, sp) : 0, kBs);
                uint64_t x = B ? bin : (~bin & ((1ull << kBs) - 1));
                r = blk * kBs + sel64(x, (uint32_t)(i - cnt));
            }
        }
        st_stream_u64(out + q, r);
    }
}

// ------------------------------------------------------------------------------------------------
// encoder
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t rrr_block_bits(uint64_t const * __restrict__ words, uint64_t nbits, uint64_t b)
{
    uint64_t pos = b * kBs;
    if (pos >= nbits)
        return 0; // the dummy block the reference appends when size % 63 == 0 (:162-164)
    uint32_t len = (nbits - pos < kBs) ? (uint32_t)(nbits - pos) : kBs;
    return read_int(words, pos, len); // the last block is zero-extended (:176-181)
}

// per block: real class and code length
__global__ void __launch_bounds__(kThreads) rrr_classify_kernel(uint64_t const * __restrict__ words,
                                                                uint64_t nbits,
                                                                uint64_t nblocks,
                                                                RrrTables const * __restrict__ tables,
                                                                uint32_t * __restrict__ blk_k,
                                                                uint32_t * __restrict__ blk_sp)
{
    uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblocks)
        return;
    uint32_t k = __popcll(rrr_block_bits(words, nbits, b));
    blk_k[b] = k;
    blk_sp[b] = (b * kBs < nbits) ? tables->space[k] : 0; // the dummy block stores nothing
}

// per superblock: the invert decision, the fused record and the three words of stored classes
__global__ void __launch_bounds__(kThreads) rrr_superblock_kernel(uint32_t const * __restrict__ blk_k,
                                                                  uint64_t const * __restrict__ ones_before,
                                                                  uint64_t const * __restrict__ bits_before,
                                                                  uint64_t nblocks,
                                                                  uint64_t nsuper,
                                                                  uint64_t * __restrict__ bt,
                                                                  uint64_t * __restrict__ records)
{
    uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g > nsuper)
        return;
    if (g == nsuper)
    { // closing record: total number of ones (m_rank's extra element, :261-262)
        records[2 * g] = ones_before[nblocks];
        records[2 * g + 1] = bits_before[nblocks];
        return;
    }
    uint64_t first = g * kK;
    bool inv = false;
    if (first + kK <= nblocks)
    { // only complete superblocks can be inverted (:203-228)
        uint32_t gt = 0;
        for (uint32_t j = 0; j < kK; ++j)
            gt += blk_k[first + j] > kBs / 2;
        inv = gt > kK / 2;
    }
    uint64_t w[3] = {0, 0, 0};
    for (uint32_t j = 0; j < kK && first + j < nblocks; ++j)
    {
        uint64_t c = blk_k[first + j];
        if (inv)
            c = kBs - c;
        uint32_t bit = j * 6;
        w[bit >> 6] |= c << (bit & 63);
        if ((bit & 63) > 58)
            w[(bit >> 6) + 1] |= c >> (64 - (bit & 63));
    }
    bt[3 * g] = w[0];
    bt[3 * g + 1] = w[1];
    bt[3 * g + 2] = w[2];
    records[2 * g] = ones_before[first];
    records[2 * g + 1] = bits_before[first] | (inv ? kInvBit : 0);
}

// per block: offset within its class (bin_to_nr, rrr_helper.hpp:346-366) OR-ed into the packed stream
__global__ void __launch_bounds__(kThreads) rrr_encode_kernel(uint64_t const * __restrict__ words,
                                                              uint64_t nbits,
                                                              uint64_t nblocks,
                                                              RrrTables const * __restrict__ tables,
                                                              uint64_t const * __restrict__ bits_before,
                                                              unsigned long long * __restrict__ btnr)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RrrTables * t = reinterpret_cast<RrrTables *>(smem_raw);
    stage_rrr(tables, t);
    uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblocks || b * kBs >= nbits)
        return;
    uint64_t bin = rrr_block_bits(words, nbits, b);
    uint32_t k = __popcll(bin);
    uint32_t sp = t->space[k];
    if (sp == 0)
        return;
    uint64_t nr = 0;
    uint32_t nn = kBs;
    while (bin)
    {
        uint32_t z = __ffsll((long long)bin) - 1; // skip zeros: they only shorten the block
        bin >>= z;
        nn -= z;
        nr += t->binom[nn - 1][k];
        --k;
        bin >>= 1;
        --nn;
    }
    uint64_t p = bits_before[b];
    uint32_t o = (uint32_t)(p & 63);
    atomicOr(btnr + (p >> 6), (unsigned long long)(nr << o));
    if (o + sp > 64)
        atomicOr(btnr + (p >> 6) + 1, (unsigned long long)(nr >> (64 - o)));
}

// select hints: one thread per superblock writes the hints whose sampled b-bit falls inside it
template <int B>
__global__ void __launch_bounds__(kThreads) rrr_hint_kernel(uint64_t const * __restrict__ records, uint64_t nsuper, uint32_t * __restrict__ hint, uint64_t nhint)
{
    uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= nsuper)
        return;
    uint64_t r0 = records[2 * g], r1 = records[2 * g + 2];
    uint64_t a = B ? r0 : g * kBs * kK - r0, e = B ? r1 : (g + 1) * kBs * kK - r1; // b-bits before / through g
    if (e <= a)
        return;
    for (uint64_t j = (a + (1ull << kHintShift) - 1) >> kHintShift; j < nhint && (j << kHintShift) + 1 <= e; ++j)
        hint[j] = (uint32_t)g;
}

int rrr_build_hints(sdslgpu_handle * h, cudaStream_t s)
{
    RrrImage & r = h->rrr;
    for (int b = 0; b < 2; ++b)
    {
        // zeros are counted over whole 2016-bit superblocks (the zero-extended tail included), like select0 does
        uint64_t args = b ? r.ones : r.nsuper * kBs * kK - r.ones;
        uint64_t nhint = args ? ((args - 1) >> kHintShift) + 1 : 0;
        SG_TRY(h->pool.alloc_t(&r.hint[b], nhint + 2));
        std::vector<uint32_t> fill(nhint + 2, (uint32_t)(r.nsuper ? r.nsuper - 1 : 0));
        SG_CUDA(cudaMemcpyAsync(r.hint[b], fill.data(), (nhint + 2) * 4, cudaMemcpyHostToDevice, s));
        SG_CUDA(cudaStreamSynchronize(s));
        if (nhint)
        {
            if (b)
                rrr_hint_kernel<1><<<blocks_for(r.nsuper), kThreads, 0, s>>>(r.records, r.nsuper, r.hint[b], nhint);
            else
                rrr_hint_kernel<0><<<blocks_for(r.nsuper), kThreads, 0, s>>>(r.records, r.nsuper, r.hint[b], nhint);
            SG_CUDA(cudaGetLastError());
        }
    }
    SG_CUDA(cudaStreamSynchronize(s));
    return SDSLGPU_OK;
}

static RrrView rrr_view(RrrImage const & r)
{
    RrrView v;
    v.size = r.size;
    v.nblocks = r.nblocks;
    v.nsuper = r.nsuper;
    v.ones = r.ones;
    v.bt = r.bt;
    v.btnr = r.btnr;
    v.records = r.records;
    v.tables = reinterpret_cast<RrrTables const *>(r.tables);
    v.hint[0] = r.hint[0];
    v.hint[1] = r.hint[1];
    return v;
}

int rrr_upload_tables(sdslgpu_handle * h, cudaStream_t s)
{
    RrrTables * d = nullptr;
    SG_TRY(h->pool.alloc_t(&d, 1));
    SG_CUDA(cudaMemcpyAsync(d, &host_tables(), sizeof(RrrTables), cudaMemcpyHostToDevice, s));
    h->rrr.tables = d;
    return SDSLGPU_OK;
}

int rrr_build(sdslgpu_handle * h, uint64_t const * words_in, bool on_device, uint64_t nbits, cudaStream_t s)
{
    RrrImage & r = h->rrr;
    r.size = nbits;
    r.nblocks = (nbits + kBs) / kBs;
    r.nsuper = (r.nblocks + kK - 1) / kK;
    uint64_t nwords = (nbits + 63) >> 6;
    SG_TRY(rrr_upload_tables(h, s));
    uint64_t * words = nullptr;
    SG_TRY(h->pool.alloc_t(&words, nwords + 2));
    SG_CUDA(cudaMemsetAsync(words + nwords, 0, 16, s));
    if (nwords)
        SG_CUDA(cudaMemcpyAsync(words, words_in, nwords * 8, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s));
    uint32_t *blk_k = nullptr, *blk_sp = nullptr;
    uint64_t *ones_before = nullptr, *bits_before = nullptr, *tmp = nullptr;
    SG_TRY(h->pool.alloc_t(&blk_k, r.nblocks));
    SG_TRY(h->pool.alloc_t(&blk_sp, r.nblocks));
    SG_TRY(h->pool.alloc_t(&ones_before, r.nblocks + 1));
    SG_TRY(h->pool.alloc_t(&bits_before, r.nblocks + 1));
    SG_TRY(h->pool.alloc_t(&tmp, scan_tmp_words(r.nblocks)));
    RrrTables const * tables = reinterpret_cast<RrrTables const *>(r.tables);
    rrr_classify_kernel<<<blocks_for(r.nblocks), kThreads, 0, s>>>(words, nbits, r.nblocks, tables, blk_k, blk_sp);
    SG_CUDA(cudaGetLastError());
    SG_CUDA(exclusive_scan(blk_k, r.nblocks, ones_before, tmp, s));
    SG_CUDA(exclusive_scan(blk_sp, r.nblocks, bits_before, tmp, s));
    uint64_t totals[2] = {0, 0};
    SG_CUDA(cudaMemcpyAsync(&totals[0], ones_before + r.nblocks, 8, cudaMemcpyDeviceToHost, s));
    SG_CUDA(cudaMemcpyAsync(&totals[1], bits_before + r.nblocks, 8, cudaMemcpyDeviceToHost, s));
    SG_CUDA(cudaStreamSynchronize(s));
    r.ones = totals[0];
    r.btnr_bits = totals[1] > 64 ? totals[1] : 64; // m_btnr has at least 64 bits (:182)
    uint64_t btnr_words = ((r.btnr_bits + 63) >> 6) + 2;
    SG_TRY(h->pool.alloc_t(&r.bt, 3 * r.nsuper + 2));
    SG_TRY(h->pool.alloc_t(&r.btnr, btnr_words));
    SG_TRY(h->pool.alloc_t(&r.records, 2 * (r.nsuper + 1) + 2));
    SG_CUDA(cudaMemsetAsync(r.btnr, 0, btnr_words * 8, s));
    SG_CUDA(cudaMemsetAsync(r.bt + 3 * r.nsuper, 0, 16, s));
    rrr_superblock_kernel<<<blocks_for(r.nsuper + 1), kThreads, 0, s>>>(blk_k, ones_before, bits_before, r.nblocks, r.nsuper, r.bt, r.records);
    SG_CUDA(cudaGetLastError());
    rrr_encode_kernel<<<blocks_for(r.nblocks), kThreads, sizeof(RrrTables), s>>>(words, nbits, r.nblocks, tables, bits_before,
                                                                                  reinterpret_cast<unsigned long long *>(r.btnr));
    SG_CUDA(cudaGetLastError());
    SG_CUDA(cudaStreamSynchronize(s));
    h->pool.release(words);
    h->pool.release(blk_k);
    h->pool.release(blk_sp);
    h->pool.release(ones_before);
    h->pool.release(bits_before);
    h->pool.release(tmp);
    return rrr_build_hints(h, s);
}

int rrr_rank_device(sdslgpu_handle const * h, int b, uint64_t const * idx, uint64_t n, uint64_t * out, cudaStream_t s)
{
    if (n == 0)
        return SDSLGPU_OK;
    rrr_rank_kernel<<<grid_for(n), kThreads, sizeof(RrrTables), s>>>(rrr_view(h->rrr), b, idx, n, out);
    SG_CUDA(cudaGetLastError());
    return SDSLGPU_OK;
}

int rrr_select_device(sdslgpu_handle const * h, int b, uint64_t const * idx, uint64_t n, uint64_t * out, cudaStream_t s)
{
    if (n == 0)
        return SDSLGPU_OK;
    if (b)
        rrr_select_kernel<1><<<grid_for(n), kThreads, sizeof(RrrTables), s>>>(rrr_view(h->rrr), idx, n, out);
    else
        rrr_select_kernel<0><<<grid_for(n), kThreads, sizeof(RrrTables), s>>>(rrr_view(h->rrr), idx, n, out);
    SG_CUDA(cudaGetLastError());
    return SDSLGPU_OK;
}

int rrr_access_device(sdslgpu_handle const * h, uint64_t const * idx, uint64_t n, uint64_t * out, cudaStream_t s)
{
    if (n == 0)
        return SDSLGPU_OK;
    rrr_access_kernel<<<grid_for(n), kThreads, sizeof(RrrTables), s>>>(rrr_view(h->rrr), idx, n, out);
    SG_CUDA(cudaGetLastError());
    return SDSLGPU_OK;
}

// SDSL-format serialisation of the device image (rrr_vector.hpp:366-378): proves the device encoder
// reproduces the reference's m_bt / m_btnr / m_btnrp / m_rank / m_invert bit for bit
int rrr_serialize(sdslgpu_handle const * h, std::vector<uint8_t> & blob)
{
    RrrImage const & r = h->rrr;
    std::vector<uint64_t> bt(3 * r.nsuper + 1), btnr((r.btnr_bits + 63) >> 6), rec(2 * (r.nsuper + 1));
    SG_CUDA(cudaMemcpy(bt.data(), r.bt, 3 * r.nsuper * 8, cudaMemcpyDeviceToHost));
    SG_CUDA(cudaMemcpy(btnr.data(), r.btnr, btnr.size() * 8, cudaMemcpyDeviceToHost));
    SG_CUDA(cudaMemcpy(rec.data(), r.records, rec.size() * 8, cudaMemcpyDeviceToHost));
    auto put64 = [&](uint64_t x) {
        for (int k = 0; k < 8; ++k)
            blob.push_back((uint8_t)(x >> (8 * k)));
    };
    auto hi = [](uint64_t x) {
        uint32_t r = 0;
        while (x >>= 1)
            ++r;
        return r;
    };
    auto put_iv = [&](uint64_t const * vals, uint64_t count, uint32_t width) { // int_vector<0> (int_vector.hpp:904-916)
        uint64_t bits = count * width;
        put64(((uint64_t)width << 56) | bits);
        std::vector<uint64_t> w(((bits + 63) >> 6) + 1, 0);
        for (uint64_t k = 0; k < count; ++k)
        {
            uint64_t p = k * width, x = vals[k] & (width >= 64 ? ~0ull : ((1ull << width) - 1));
            w[p >> 6] |= x << (p & 63);
            if ((p & 63) + width > 64)
                w[(p >> 6) + 1] |= x >> (64 - (p & 63));
        }
        for (uint64_t k = 0; k < ((bits + 63) >> 6); ++k)
            put64(w[k]);
    };
    put64(r.size);
    // m_bt: nblocks x 6 bits — the device keeps exactly these words
    {
        uint64_t bits = r.nblocks * 6;
        put64((6ull << 56) | bits);
        for (uint64_t k = 0; k < ((bits + 63) >> 6); ++k)
            put64(bt[k]);
    }
    put64((1ull << 56) | r.btnr_bits);
    for (uint64_t k = 0; k < btnr.size(); ++k)
        put64(btnr[k]);
    uint64_t total_bits = rec[2 * r.nsuper + 1];
    std::vector<uint64_t> p(r.nsuper), rk, inv(r.nsuper);
    for (uint64_t g = 0; g < r.nsuper; ++g)
    {
        p[g] = rec[2 * g + 1] & ~kInvBit;
        inv[g] = (rec[2 * g + 1] & kInvBit) ? 1 : 0;
        rk.push_back(rec[2 * g]);
    }
    // a trailing superblock that only holds the dummy block keeps btnrp == 0 in the reference (:240-258)
    if (r.nsuper && (r.nsuper - 1) * kK * kBs >= r.size && r.size > 0)
        p[r.nsuper - 1] = 0;
    if (r.size % (kK * kBs))
        rk.push_back(rec[2 * r.nsuper]);
    else if (!rk.empty())
        rk.back() = rec[2 * r.nsuper]; // m_rank[last] = total (:261-262)
    put_iv(p.data(), p.size(), hi(total_bits) + 1);
    put_iv(rk.data(), rk.size(), hi(r.ones) + 1);
    put_iv(inv.data(), inv.size(), 1);
    return SDSLGPU_OK;
}

} // namespace sdslgpu
