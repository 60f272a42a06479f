// rrr.cu — rrr_vector<63, int_vector<>, 32>: device-side encoder and batched rank / select / access.
//
// Replaces (results bit-exact; the encoding is bit-identical to the reference's, so SDSL-serialised
// vectors are ingested without re-encoding and the device image serialises back to the same bytes):
//   rrr_vector ctor            rrr_vector.hpp:158-270 + rrr_helper.hpp:346-366 (bin_to_nr) -> rrr_classify /
//                                                                  rrr_superblock / rrr_encode kernels + scans
//   rank_support_rrr::rank     rrr_vector.hpp:503-544  -> rrr_rank_kernel
//   select_support_rrr::select rrr_vector.hpp:639-726  -> rrr_select_kernel<B>
//   rrr_vector::operator[]     rrr_vector.hpp:276-298  -> rrr_access_kernel
//
// Device layout.  The reference keeps five arrays (m_bt, m_btnr, m_btnrp, m_rank, m_invert) and a query
// touches four of them at unrelated addresses.  A B200 pays per cache line touched, so all
// per-superblock metadata is fused into ONE 64-byte record (half a line, two LDG.256):
//     w0      ones before the superblock                         (m_rank[g])
//     w1      bit offset into the offset stream | invert << 63   (m_btnrp[g], m_invert[g])
//     w2..w4  the 32 six-bit classes exactly as stored in m_bt   (192 bits)
//     w5      prefix sums after 8 / 16 / 24 blocks: ones (10+10+11 bits) and offset bits (10+10+11 bits)
//     w6      ones in the superblock, w7 offset bits of the superblock
// so rank = record + one read of the offset stream (m_btnr, unchanged), and the class scan is at most 7 steps.
// C(n,k) for n <= 62 (23.6 KB: 64-bit rows for n >= 34, 32-bit rows below) and the code lengths live in shared memory.
#include "binned.cuh"
#include "internal.h"
#include "rrr_device.cuh"
#include "scan.cuh"

namespace sdslgpu
{

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
        uint64_t res = SDSLGPU_NPOS;
        if (i < v.size)
        {
            uint64_t blk = i / kBs, g = blk / kK;
            uint32_t off = (uint32_t)(i - blk * kBs), nblk = (uint32_t)(blk - g * kK);
            RrrRecord r;
            ld_record(v.records, g, r);
            bool inv = (r.w[1] & kInvBit) != 0;
            uint32_t k = rec_class(r.w[2], r.w[3], r.w[4], nblk);
            if (inv)
                k = kBs - k;
            if (k == 0 || k == kBs)
                res = k != 0; // rrr_vector.hpp:283-288
            else
            {
                uint64_t p = r.w[1] & ~kInvBit, ones = 0;
                rec_prefix(r, t, nblk, inv, ones, p);
                uint32_t bit;
                rrr_prefix_ones(t, k, read_int(v.btnr, p, t->space[k]), off, true, bit, v.try_sparse != 0);
                res = bit;
            }
        }
        st_stream_u64(out + q, res);
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
        uint64_t res;
        if (i == 0)
            res = SDSLGPU_NPOS;
        else if (i > args)
            res = v.size; // the reference's in-band answer (rrr_vector.hpp:641-642, 686-689)
        else
            res = rrr_select_one<B>(v, t, i);
        st_stream_u64(out + q, res);
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

__global__ void __launch_bounds__(kThreads) rrr_superblock_kernel(uint32_t const * __restrict__ blk_k,
                                                                  uint32_t const * __restrict__ blk_sp,
                                                                  uint64_t const * __restrict__ ones_before,
                                                                  uint64_t const * __restrict__ bits_before,
                                                                  uint64_t nbits,
                                                                  uint64_t nblocks,
                                                                  uint64_t nsuper,
                                                                  RrrTables const * __restrict__ tables,
                                                                  uint64_t * __restrict__ records)
{
    uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g > nsuper)
        return;
    uint64_t * rec = records + g * kRecWords;
    if (g == nsuper)
    { // closing record: totals (m_rank's extra element, rrr_vector.hpp:261-262)
        rec[0] = ones_before[nblocks];
        rec[1] = bits_before[nblocks];
        for (uint32_t k = 2; k < kRecWords; ++k)
            rec[k] = 0;
        return;
    }
    uint64_t first = g * kK;
    uint32_t k_real[kK];
    uint32_t here = 0;
    for (uint32_t j = 0; j < kK; ++j)
    {
        k_real[j] = (first + j < nblocks) ? blk_k[first + j] : 0;
        // blocks that start at or past nbits (the dummy block) carry a class but no offset bits
        here += (first + j < nblocks);
    }
    // the dummy block must not contribute offset bits: its real class is 0 => space[0] == 0 already
    (void)blk_sp;
    (void)nbits;
    rrr_make_record(k_real, here, first + kK <= nblocks, tables->space, ones_before[first], bits_before[first], rec);
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
        nr += rrr_binom(t, nn - 1, k);
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
__global__ void __launch_bounds__(kThreads) rrr_hint_kernel(uint64_t const * __restrict__ records, uint64_t nsuper, uint32_t * __restrict__ hint, uint64_t nhint, uint32_t shift)
{
    uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= nsuper)
        return;
    uint64_t r0 = records[g * kRecWords], r1 = records[(g + 1) * kRecWords];
    uint64_t a = B ? r0 : g * kBs * kK - r0, e = B ? r1 : (g + 1) * kBs * kK - r1; // b-bits before / through g
    if (e <= a)
        return;
    for (uint64_t j = (a + (1ull << shift) - 1) >> shift; j < nhint && (j << shift) + 1 <= e; ++j)
        hint[j] = (uint32_t)g;
}

int rrr_build_hints(DevicePool & pool, RrrImage & r, cudaStream_t s)
{
    for (int b = 0; b < 2; ++b)
    {
        // zeros are counted over whole 2016-bit superblocks (the zero-extended tail included), like select0 does
        uint64_t args = b ? r.ones : r.nsuper * kBs * kK - r.ones;
        uint32_t const shift = rrr_hint_shift(args, r.nsuper);
        r.hint_shift[b] = shift;
        uint64_t nhint = args ? ((args - 1) >> shift) + 1 : 0;
        SG_TRY(pool.alloc_t(&r.hint[b], nhint + 2));
        std::vector<uint32_t> fill(nhint + 2, (uint32_t)(r.nsuper ? r.nsuper - 1 : 0));
        SG_CUDA(cudaMemcpyAsync(r.hint[b], fill.data(), (nhint + 2) * 4, cudaMemcpyHostToDevice, s));
        SG_CUDA(cudaStreamSynchronize(s));
        if (nhint)
        {
            if (b)
                rrr_hint_kernel<1><<<blocks_for(r.nsuper), kThreads, 0, s>>>(r.records, r.nsuper, r.hint[b], nhint, shift);
            else
                rrr_hint_kernel<0><<<blocks_for(r.nsuper), kThreads, 0, s>>>(r.records, r.nsuper, r.hint[b], nhint, shift);
            SG_CUDA(cudaGetLastError());
        }
    }
    SG_CUDA(cudaStreamSynchronize(s));
    return SDSLGPU_OK;
}

int rrr_upload_tables(DevicePool & pool, RrrImage & r, cudaStream_t s)
{
    RrrTables * d = nullptr;
    SG_TRY(pool.alloc_t(&d, 1));
    SG_CUDA(cudaMemcpyAsync(d, &host_tables(), sizeof(RrrTables), cudaMemcpyHostToDevice, s));
    r.tables = d;
    return SDSLGPU_OK;
}

int rrr_build_image(DevicePool & pool, RrrImage & r, uint64_t const * words_in, bool on_device, uint64_t nbits, cudaStream_t s)
{
    r.size = nbits;
    r.nblocks = (nbits + kBs) / kBs;
    r.nsuper = (r.nblocks + kK - 1) / kK;
    uint64_t nwords = (nbits + 63) >> 6;
    SG_TRY(rrr_upload_tables(pool, r, s));
    uint64_t * words = nullptr;
    SG_TRY(pool.alloc_t(&words, nwords + 2));
    SG_CUDA(cudaMemsetAsync(words + nwords, 0, 16, s));
    if (nwords)
        SG_CUDA(cudaMemcpyAsync(words, words_in, nwords * 8, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s));
    uint32_t *blk_k = nullptr, *blk_sp = nullptr;
    uint64_t *ones_before = nullptr, *bits_before = nullptr, *tmp = nullptr;
    SG_TRY(pool.alloc_t(&blk_k, r.nblocks));
    SG_TRY(pool.alloc_t(&blk_sp, r.nblocks));
    SG_TRY(pool.alloc_t(&ones_before, r.nblocks + 1));
    SG_TRY(pool.alloc_t(&bits_before, r.nblocks + 1));
    SG_TRY(pool.alloc_t(&tmp, scan_tmp_words(r.nblocks)));
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
    SG_TRY(pool.alloc_t(&r.btnr, btnr_words));
    SG_TRY(pool.alloc_t(&r.records, kRecWords * (r.nsuper + 1)));
    SG_CUDA(cudaMemsetAsync(r.btnr, 0, btnr_words * 8, s));
    rrr_superblock_kernel<<<blocks_for(r.nsuper + 1), kThreads, 0, s>>>(blk_k, blk_sp, ones_before, bits_before, nbits, r.nblocks, r.nsuper, tables, r.records);
    SG_CUDA(cudaGetLastError());
    rrr_encode_kernel<<<blocks_for(r.nblocks), kThreads, sizeof(RrrTables), s>>>(words, nbits, r.nblocks, tables, bits_before,
                                                                                  reinterpret_cast<unsigned long long *>(r.btnr));
    SG_CUDA(cudaGetLastError());
    SG_CUDA(cudaStreamSynchronize(s));
    pool.release(words);
    pool.release(blk_k);
    pool.release(blk_sp);
    pool.release(ones_before);
    pool.release(bits_before);
    pool.release(tmp);
    return rrr_build_hints(pool, r, s);
}

int rrr_build(sdslgpu_handle * h, uint64_t const * words_in, bool on_device, uint64_t nbits, cudaStream_t s)
{
    return rrr_build_image(h->pool, h->rrr, words_in, on_device, nbits, s);
}

// records from the reference's own arrays (ingest of a serialised rrr_vector): stored classes + samples
int rrr_records_from_sdsl(DevicePool & pool, RrrImage & r,
                          uint64_t const * bt_words /* packed 6-bit stored classes */,
                          uint64_t nblocks,
                          std::vector<uint64_t> const & rank,
                          std::vector<uint64_t> const & btnrp,
                          std::vector<uint8_t> const & invert,
                          uint64_t total_bits_hint,
                          cudaStream_t s)
{
    std::vector<uint64_t> rec;
    rrr_records_host(bt_words, nblocks, r.nsuper, r.ones, rank, btnrp, invert, total_bits_hint, rec);
    SG_TRY(pool.alloc_t(&r.records, rec.size()));
    SG_CUDA(cudaMemcpyAsync(r.records, rec.data(), rec.size() * 8, cudaMemcpyHostToDevice, s));
    SG_CUDA(cudaStreamSynchronize(s));
    return SDSLGPU_OK;
}

// op of the locality-ordered batch pipeline (binned.cuh): rank(i) reads the 64-byte record of superblock i / 2016
// and the offset bits right behind btnrp — both grow with i
#ifndef BIN_RRR_CTAS
#define BIN_RRR_CTAS 6 // resident CTAs per SM the rrr ops are compiled for (tools/variants.sh measures 4 / 6 / 8)
#endif
struct RrrRankOp
{
    static constexpr int kIlp = 1;
    static constexpr int kMinCtas = BIN_RRR_CTAS;
    static constexpr uint32_t kSmem = sizeof(RrrTables);
    RrrView v;
    int b;
    RrrTables const * t;
    __device__ __forceinline__ void stage(uint8_t * smem)
    {
        RrrTables * st = reinterpret_cast<RrrTables *>(smem);
        stage_rrr(v.tables, st);
        t = st;
    }
    __device__ __forceinline__ uint64_t operator()(uint64_t i) const
    {
        uint64_t r = rrr_rank1_one(v, t, i);
        return b ? r : i - r;
    }
};

// select(i): hint -> bisection over the superblock records -> class scan -> decode, all next to the i-th b-bit.
// key = i - 1; key == #b-bits stands for every i past the end (the reference's in-band size(), rrr_vector.hpp:641-642)
template <int B>
struct RrrSelectOp
{
    static constexpr int kIlp = 1;
    static constexpr int kMinCtas = BIN_RRR_CTAS;
    static constexpr uint32_t kSmem = sizeof(RrrTables);
    RrrView v;
    uint64_t args;
    RrrTables const * t;
    __device__ __forceinline__ void stage(uint8_t * smem)
    {
        RrrTables * st = reinterpret_cast<RrrTables *>(smem);
        stage_rrr(v.tables, st);
        t = st;
    }
    __device__ __forceinline__ uint64_t operator()(uint64_t key) const
    {
        return key >= args ? v.size : rrr_select_one<B>(v, t, key + 1);
    }
};

static uint64_t rrr_index_bytes(RrrImage const & r)
{
    return (r.nsuper + 1) * 64 + ((r.btnr_bits + 63) >> 6) * 8;
}

int rrr_rank_image(RrrImage const & r, int b, uint64_t const * idx, uint64_t n, uint64_t * out, cudaStream_t s, int order)
{
    if (n == 0)
        return SDSLGPU_OK;
    uint64_t index_bytes = rrr_index_bytes(r);
    if (bin_wanted(order, index_bytes, n))
    {
        bool done = false;
        SG_TRY(bin_run(RrrRankOp{rrr_view(r), b, nullptr}, index_bytes, 0, r.size, idx, n, out, s, &done));
        if (done)
            return SDSLGPU_OK;
    }
    rrr_rank_kernel<<<grid_for(n), kThreads, sizeof(RrrTables), s>>>(rrr_view(r), b, idx, n, out);
    SG_CUDA(cudaGetLastError());
    return SDSLGPU_OK;
}

int rrr_rank_device(sdslgpu_handle const * h, int b, uint64_t const * idx, uint64_t n, uint64_t * out, cudaStream_t s)
{
    return rrr_rank_image(h->rrr, b, idx, n, out, s, h->order);
}

int rrr_select_device(sdslgpu_handle const * h, int b, uint64_t const * idx, uint64_t n, uint64_t * out, cudaStream_t s)
{
    if (n == 0)
        return SDSLGPU_OK;
    if (bin_wanted(h->order, rrr_index_bytes(h->rrr), n))
    {
        RrrImage const & r = h->rrr;
        uint64_t args = b ? r.ones : r.size - r.ones;
        bool done = false;
        if (b)
            SG_TRY(bin_run(RrrSelectOp<1>{rrr_view(r), args, nullptr}, rrr_index_bytes(r), 1, args, idx, n, out, s, &done, true));
        else
            SG_TRY(bin_run(RrrSelectOp<0>{rrr_view(r), args, nullptr}, rrr_index_bytes(r), 1, args, idx, n, out, s, &done, true));
        if (done)
            return SDSLGPU_OK;
    }
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
    return rrr_serialize_image(h->rrr, blob);
}

int rrr_serialize_image(RrrImage const & r, std::vector<uint8_t> & blob)
{
    std::vector<uint64_t> btnr((r.btnr_bits + 63) >> 6), rec(kRecWords * (r.nsuper + 1));
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
    // m_bt: nblocks x 6 bits = the class words of the records, 3 per superblock
    {
        uint64_t bits = r.nblocks * 6, nw = (bits + 63) >> 6;
        put64((6ull << 56) | bits);
        for (uint64_t k = 0; k < nw; ++k)
            put64(rec[(k / 3) * kRecWords + 2 + (k % 3)]);
    }
    put64((1ull << 56) | r.btnr_bits);
    for (uint64_t k = 0; k < btnr.size(); ++k)
        put64(btnr[k]);
    uint64_t total_bits = rec[kRecWords * r.nsuper + 1];
    std::vector<uint64_t> p(r.nsuper), rk, inv(r.nsuper);
    for (uint64_t g = 0; g < r.nsuper; ++g)
    {
        p[g] = rec[g * kRecWords + 1] & ~kInvBit;
        inv[g] = (rec[g * kRecWords + 1] & kInvBit) ? 1 : 0;
        rk.push_back(rec[g * kRecWords]);
    }
    // a trailing superblock that only holds the dummy block keeps btnrp == 0 in the reference (:240-258)
    if (r.nsuper && (r.nsuper - 1) * kK * kBs >= r.size && r.size > 0)
        p[r.nsuper - 1] = 0;
    if (r.size % (kK * kBs))
        rk.push_back(rec[kRecWords * r.nsuper]);
    else if (!rk.empty())
        rk.back() = rec[kRecWords * r.nsuper]; // m_rank[last] = total (:261-262)
    put_iv(p.data(), p.size(), hi(total_bits) + 1);
    put_iv(rk.data(), rk.size(), hi(r.ones) + 1);
    put_iv(inv.data(), inv.size(), 1);
    return SDSLGPU_OK;
}

} // namespace sdslgpu
