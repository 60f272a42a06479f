// bv.cu — plain bit_vector: device builders and batched rank / select / access kernels.
//
// Replaces (results bit-exact, layout re-designed for B200 — DESIGN.md §3):
//   rank_support_v<b,1>  ctor  rank_support_v.hpp:72-122     -> bv_pack_kernel + scan + bv_finish_kernel
//   rank_support_v<b,1>::rank  rank_support_v.hpp:129-139    -> bv_rank_kernel<B>   (1 sector / query)
//   select_support_mcl<b,1> ctor select_support_mcl.hpp:207-381 -> bv_samples_kernel<B>
//   select_support_mcl<b,1>::select  :384-439                -> bv_select_kernel<B>
//   bit_vector::operator[]     int_vector.hpp:1900-1904      -> bv_access_kernel
// plus, under SDSLGPU_F_SDSL_LAYOUT, the reference's own table (m_basic_block) built on the device,
// byte-identical to the reference's, and a rank kernel that reads it exactly as the reference does.
#include <cstdlib>

#include "fan.cuh"
#include "internal.h"
#include "bv_device.cuh"
#include "scan.cuh"

namespace sdslgpu
{

// ------------------------------------------------------------------------------------------------
// build, phase 1: re-chunk the LSB-first bit stream into 224-bit payloads and count each block
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
    bv_pack_kernel(uint32_t const * __restrict__ w32, uint64_t nbits, uint64_t nblocks, bvblock * __restrict__ blocks, uint32_t * __restrict__ blk_ones)
{
    uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblocks)
        return;
    uint64_t n32 = (nbits + 31) >> 5; // number of 32-bit words holding valid bits
    uint32_t d[7];
    uint32_t c = 0;
#pragma unroll
    for (int j = 0; j < 7; ++j)
    {
        uint64_t k = b * 7 + j;
        uint32_t x = 0;
        if (k < n32)
        {
            x = w32[k];
            if (k == n32 - 1 && (nbits & 31))
                x &= (1u << (nbits & 31)) - 1u; // bits past nbits are unspecified in SDSL: ignore them
        }
        d[j] = x;
        c += __popc(x);
    }
    // two 128-bit stores = one full sector
    uint4 * o = reinterpret_cast<uint4 *>(blocks + b);
    o[0] = make_uint4(0u, d[0], d[1], d[2]);
    o[1] = make_uint4(d[3], d[4], d[5], d[6]);
    blk_ones[b] = c;
}

// build, phase 3: turn absolute prefix counts into (top[superblock], 32-bit in-superblock count)
__global__ void __launch_bounds__(kThreads)
    bv_finish_kernel(uint64_t const * __restrict__ abs_ones, uint64_t nblocks, bvblock * __restrict__ blocks, uint64_t * __restrict__ top)
{
    uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblocks)
        return;
    uint64_t sb = b >> kSuperShift;
    uint64_t base = abs_ones[sb << kSuperShift];
    blocks[b].cnt = (uint32_t)(abs_ones[b] - base);
    if ((b & ((1ull << kSuperShift) - 1)) == 0)
        top[sb] = base;
}

// select samples: samp[j] = block that holds the (j*S+1)-th B-bit.  One thread per block writes the
// (at most 224/S + 1) samples that fall inside it.
template <int B>
__global__ void __launch_bounds__(kThreads) bv_samples_kernel(uint64_t const * __restrict__ abs_ones,
                                                              uint32_t const * __restrict__ blk_ones,
                                                              uint64_t nbits,
                                                              uint64_t nblocks,
                                                              uint32_t log_s,
                                                              uint32_t * __restrict__ samp,
                                                              uint64_t nsamp,
                                                              bvblock const * __restrict__ blocks,
                                                              uint32_t pos_mode)
{
    uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblocks)
        return;
    uint64_t first = b * kBlockBits;
    uint64_t valid = (first >= nbits) ? 0 : ((nbits - first < kBlockBits) ? nbits - first : kBlockBits);
    uint64_t a1 = abs_ones[b], c1 = blk_ones[b];
    uint64_t a = B ? a1 : first - a1; // B-bits before the block (first <= nbits here whenever valid > 0)
    uint64_t c = B ? c1 : valid - c1;
    if (c == 0)
        return;
    uint64_t S = 1ull << log_s;
    uint32_t cnt = 0, d[7] = {0, 0, 0, 0, 0, 0, 0};
    bool loaded = false;
    for (uint64_t j = (a + S - 1) >> log_s; j < nsamp && (j << log_s) + 1 <= a + c; ++j)
    {
        if (!pos_mode)
        {
            samp[j] = (uint32_t)b;
            continue;
        }
        if (!loaded)
        {
            ld_block(blocks + b, cnt, d);
            loaded = true;
        }
        // the sampled bit's position in 32-bit chunks (its block is chunk / 7)
        samp[j] = (uint32_t)((first + block_select<B>(d, (uint32_t)((j << log_s) + 1 - a))) >> 5);
    }
}

// select sectors (bv_device.cuh): one thread per sector
template <int B>
__global__ void __launch_bounds__(kThreads) bv_select_sectors_kernel(BvView const v, uint64_t nblocks, uint64_t args, uint32_t stride, uint64_t nsect, bvblock * __restrict__ sect)
{
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nsect)
        return;
    uint32_t c0, d[7];
    bv_make_sector<B>(v, nblocks, args, stride, j, c0, d);
    uint4 * o = reinterpret_cast<uint4 *>(sect + j);
    o[0] = make_uint4(c0, d[0], d[1], d[2]);
    o[1] = make_uint4(d[3], d[4], d[5], d[6]);
}

// ------------------------------------------------------------------------------------------------
// rank: one 32-byte sector per query
// ------------------------------------------------------------------------------------------------
// kFan (multi-GPU group calls, group.cu; fan.cuh): 1 = every result is also stored to the same index of the other
// members' arrays, 2 = as packed fields into their staging regions.  The loop bound is warp-uniform (all lanes of a warp
// leave together) so that the packed form can shuffle.
template <int B, int ILP, int kFan>
__global__ void __launch_bounds__(kThreads) bv_rank_kernel(bvblock const * __restrict__ blocks,
                                                           uint64_t const * __restrict__ top,
                                                           uint64_t nbits,
                                                           uint64_t const * __restrict__ idx,
                                                           uint64_t n,
                                                           uint64_t * __restrict__ out,
                                                           Fan const fan)
{
    uint64_t const stride = (uint64_t)gridDim.x * blockDim.x * ILP;
    uint32_t const lane = threadIdx.x & 31u;
    for (uint64_t wbase = (uint64_t)blockIdx.x * blockDim.x * ILP + (threadIdx.x - lane); wbase < n; wbase += stride)
    {
        uint64_t const base = wbase + lane;
        uint64_t i[ILP], blk[ILP];
        uint32_t cnt[ILP], d[ILP][7];
        bool ok[ILP];
#pragma unroll
        for (int u = 0; u < ILP; ++u)
        {
            uint64_t q = base + (uint64_t)u * blockDim.x;
            i[u] = (q < n) ? ld_stream_u64(idx + q) : 0;
            ok[u] = i[u] <= nbits;
            blk[u] = ok[u] ? i[u] / kBlockBits : 0;
        }
#pragma unroll
        for (int u = 0; u < ILP; ++u)
            ld_block_half_line(blocks + blk[u], cnt[u], d[u]);
#pragma unroll
        for (int u = 0; u < ILP; ++u)
        {
            uint64_t q = base + (uint64_t)u * blockDim.x;
            uint32_t rem = (uint32_t)(i[u] - blk[u] * kBlockBits);
            uint64_t r = __ldg(top + (blk[u] >> kSuperShift)) + cnt[u] + block_prefix_popc(d[u], rem);
            if (!B)
                r = i[u] - r;
            r = ok[u] ? r : SDSLGPU_NPOS;
            if (q < n)
            {
                st_stream_u64(out + q, r);
                if (kFan == 1)
                    fan_store(fan, q, r);
            }
            if (kFan == 2)
            {
                uint64_t const q0 = q - lane;
                if (q0 < n)
                    fan_store_packed(fan, q0, lane, n - q0 < 32 ? (uint32_t)(n - q0) : 32u, r);
            }
        }
    }
}

// rank on the reference's own layout (SDSLGPU_F_SDSL_LAYOUT): 16-byte table pair + 8-byte data word,
// i.e. the two gathers of rank_support_v.hpp:133-135 issued together.
template <int B, int ILP>
__global__ void __launch_bounds__(kThreads) bv_rank_sdsl_kernel(uint64_t const * __restrict__ words,
                                                                uint64_t const * __restrict__ table,
                                                                uint64_t nbits,
                                                                uint64_t const * __restrict__ idx,
                                                                uint64_t n,
                                                                uint64_t * __restrict__ out)
{
    uint64_t const stride = (uint64_t)gridDim.x * blockDim.x * ILP;
    for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x * ILP + threadIdx.x; base < n; base += stride)
    {
        uint64_t i[ILP], a[ILP], r9[ILP], w[ILP];
        bool ok[ILP];
#pragma unroll
        for (int u = 0; u < ILP; ++u)
        {
            uint64_t q = base + (uint64_t)u * blockDim.x;
            i[u] = (q < n) ? ld_stream_u64(idx + q) : 0;
            ok[u] = i[u] <= nbits;
            if (!ok[u])
                i[u] = 0;
        }
#pragma unroll
        for (int u = 0; u < ILP; ++u)
        {
            ld_pair(table + ((i[u] >> 8) & ~1ULL), a[u], r9[u]);
            w[u] = ld_nc_u64(words + (i[u] >> 6)); // the pad word makes i == nbits safe
        }
#pragma unroll
        for (int u = 0; u < ILP; ++u)
        {
            uint64_t q = base + (uint64_t)u * blockDim.x;
            if (q < n)
            {
                uint64_t x = B ? w[u] : ~w[u];
                uint64_t r = a[u] + ((r9[u] >> (63 - 9 * ((i[u] & 0x1FF) >> 6))) & 0x1FF) + __popcll(x & lo_set64((uint32_t)(i[u] & 63)));
                st_stream_u64(out + q, ok[u] ? r : SDSLGPU_NPOS);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// select
// ------------------------------------------------------------------------------------------------
template <int B, int kFan>
__global__ void __launch_bounds__(kThreads)
    bv_select_kernel(BvView const v, uint64_t args, uint64_t const * __restrict__ idx, uint64_t n, uint64_t * __restrict__ out, Fan const fan)
{
    uint64_t const stride = (uint64_t)gridDim.x * blockDim.x;
    uint32_t const lane = threadIdx.x & 31u;
    for (uint64_t q0 = (uint64_t)blockIdx.x * blockDim.x + (threadIdx.x - lane); q0 < n; q0 += stride)
    { // warp-uniform bound: the packed fan-out shuffles
        uint64_t const q = q0 + lane;
        uint64_t r = SDSLGPU_NPOS;
        if (q < n)
        {
            uint64_t i = ld_stream_u64(idx + q);
            if (i >= 1 && i <= args)
            {
                r = bv_select_any<B>(v, i); // select sectors, once a large batch has had them built
            }
            st_stream_u64(out + q, r);
            if (kFan == 1)
                fan_store(fan, q, r);
        }
        if (kFan == 2)
            fan_store_packed(fan, q0, lane, n - q0 < 32 ? (uint32_t)(n - q0) : 32u, r);
    }
}

__global__ void __launch_bounds__(kThreads)
    bv_access_kernel(bvblock const * __restrict__ blocks, uint64_t nbits, uint64_t const * __restrict__ idx, uint64_t n, uint64_t * __restrict__ out)
{
    uint64_t const stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += stride)
    {
        uint64_t i = ld_stream_u64(idx + q);
        uint64_t r = SDSLGPU_NPOS;
        if (i < nbits)
        {
            uint64_t b = i / kBlockBits;
            uint32_t rem = (uint32_t)(i - b * kBlockBits);
            r = (ld_nc_u32(&blocks[b].d[rem >> 5]) >> (rem & 31)) & 1u;
        }
        st_stream_u64(out + q, r);
    }
}

// ------------------------------------------------------------------------------------------------
// the reference's m_basic_block, built on the device (rank_support_v.hpp:72-122), byte-identical.
// One thread per 512-bit superblock: 8 word popcounts -> 7 nine-bit prefixes + the superblock total.
// ------------------------------------------------------------------------------------------------
template <int B>
__global__ void __launch_bounds__(kThreads)
    sdsl_table_rel_kernel(uint64_t const * __restrict__ words, uint64_t nwords, uint64_t nsuper, uint64_t * __restrict__ table, uint32_t * __restrict__ sb_total)
{
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nsuper)
        return;
    uint64_t second = 0, sum = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j)
    {
        uint64_t wi = k * 8 + j;
        if (j > 0 && wi <= nwords) // the reference also records the prefix just past the last word (:109-113)
            second |= sum << (63 - 9 * j);
        if (wi < nwords)
        {
            uint64_t x = words[wi];
            sum += __popcll(B ? x : ~x);
        }
    }
    table[2 * k + 1] = second;
    sb_total[k] = (uint32_t)sum;
}

// rank_support_v5's table (rank_support_v5.hpp:66-122): 2048-bit superblocks; the second word of a pair packs the
// prefix counts after 6, 12, 18, 24 and 30 words of the superblock into 12-bit fields at bits 48, 36, 24, 12, 0
// (the prefix just past the last word is recorded too when it falls on a multiple of 6 words).
template <int B>
__global__ void __launch_bounds__(kThreads)
    sdsl_table5_rel_kernel(uint64_t const * __restrict__ words, uint64_t nwords, uint64_t nsuper, uint64_t * __restrict__ table, uint32_t * __restrict__ sb_total)
{
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nsuper)
        return;
    uint64_t second = 0, sum = 0;
    for (int j = 0; j < 32; ++j)
    {
        uint64_t wi = k * 32 + j;
        if (j > 0 && j % 6 == 0 && wi <= nwords)
            second |= sum << (60 - 12 * (j / 6));
        if (wi < nwords)
        {
            uint64_t x = words[wi];
            sum += __popcll(B ? x : ~x);
        }
    }
    table[2 * k + 1] = second;
    sb_total[k] = (uint32_t)sum;
}

__global__ void __launch_bounds__(kThreads) sdsl_table_abs_kernel(uint64_t const * __restrict__ abs, uint64_t nsuper, uint64_t * __restrict__ table)
{
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < nsuper)
        table[2 * k] = abs[k];
}

unsigned grid_for(uint64_t n, int per_thread)
{
    uint64_t want = (n + (uint64_t)kThreads * per_thread - 1) / ((uint64_t)kThreads * per_thread);
    uint64_t cap = (uint64_t)sm_count() * 8; // 8 resident CTAs of 256 threads per SM = 64 warps
    if (want < 1)
        want = 1;
    return (unsigned)(want < cap ? want : cap);
}
unsigned blocks_for(uint64_t n)
{
    return (unsigned)((n + kThreads - 1) / kThreads);
}

int bv_build(DevicePool & pool, BvImage & v, uint32_t flags, uint64_t const * words_in, bool on_device, uint64_t nbits, cudaStream_t s)
{
    v.nbits = nbits;
    v.nwords = (nbits + 63) >> 6;
    v.nblocks = nbits / kBlockBits + 1;
    if (v.nblocks >= (1ull << 32))
    {
        set_error("bit vector too large: %llu bits (limit 2^32 blocks of 224 bits)", (unsigned long long)nbits);
        return SDSLGPU_EINVAL;
    }
    v.ntop = ((v.nblocks - 1) >> kSuperShift) + 1;

    // raw words on the device (+1 zero pad word, like SDSL's allocation, memory_management.hpp:901-906)
    uint64_t * words = nullptr;
    SG_TRY(pool.alloc_t(&words, v.nwords + 2));
    SG_CUDA(cudaMemsetAsync(words + v.nwords, 0, 16, s));
    if (v.nwords)
        SG_CUDA(cudaMemcpyAsync(words, words_in, v.nwords * 8, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s));

    uint32_t * blk_ones = nullptr;
    uint64_t * abs_ones = nullptr;
    uint64_t * tmp = nullptr;
    SG_TRY(pool.alloc_t(&v.blocks, v.nblocks));
    SG_TRY(pool.alloc_t(&v.top, v.ntop));
    SG_TRY(pool.alloc_t(&blk_ones, v.nblocks));
    SG_TRY(pool.alloc_t(&abs_ones, v.nblocks + 1));
    SG_TRY(pool.alloc_t(&tmp, scan_tmp_words(v.nblocks)));

    bv_pack_kernel<<<blocks_for(v.nblocks), kThreads, 0, s>>>(reinterpret_cast<uint32_t const *>(words), nbits, v.nblocks, v.blocks, blk_ones);
    SG_CUDA(cudaGetLastError());
    SG_CUDA(exclusive_scan(blk_ones, v.nblocks, abs_ones, tmp, s));
    bv_finish_kernel<<<blocks_for(v.nblocks), kThreads, 0, s>>>(abs_ones, v.nblocks, v.blocks, v.top);
    SG_CUDA(cudaGetLastError());
    SG_CUDA(cudaMemcpyAsync(&v.ones, abs_ones + v.nblocks, 8, cudaMemcpyDeviceToHost, s));
    SG_CUDA(cudaStreamSynchronize(s));

    if (!(flags & SDSLGPU_F_NO_SELECT))
    {
        for (int b = 0; b < 2; ++b)
        {
            uint64_t m = b ? v.ones : nbits - v.ones;
            // sample stride S = 2^log_s.  Small vectors: the largest power of two <= 64 with S <= 128 * density
            // (samples ~128 bits apart, the hinted block is almost always the right one).  Large vectors: S grows
            // (up to 4096) until the u32 sample table is <= 48 MB, so the table stays resident in the 126 MB L2 and
            // a query pays ONE DRAM line (the sector block found by interpolating between two samples) instead of two.
            // (48 MB, not 32: a 2^33-bit vector of density 1/2 has 2^32 +- 50 K ones — a 32 MB limit put S = 512 and
            // S = 1024 on either side of that coin flip, and S = 1024 costs 12 % in both batch orders.)
            uint32_t ls = 6;
            while (ls > 0 && ((1ull << ls) * nbits > 128ull * m * 1ull) && m > 0)
                --ls;
            while (ls < 12 && 4ull * (m >> ls) > (48ull << 20))
                ++ls;
            if (char const * e = std::getenv("SDSLGPU_SELECT_LOG_S")) // tuning knob for experiments
                ls = (uint32_t)std::atoi(e) > 16 ? 16u : (uint32_t)std::atoi(e);
            v.log_s[b] = ls;
            v.interp[b] = ls > 6;
            if (char const * e = std::getenv("SDSLGPU_SELECT_INTERP"))
                v.interp[b] = std::atoi(e) != 0;
            v.nsamp[b] = m ? ((m - 1) >> ls) + 1 : 0;
            SG_TRY(pool.alloc_t(&v.samp[b], v.nsamp[b] + 2));
            // position-valued samples (bv_device.cuh bv_select) while positions >> 5 fit 32 bits
            v.samp_pos[b] = nbits <= (1ull << 36); // chunk indices (and the + 6 sentinel) stay below 2^32
            if (char const * e = std::getenv("SDSLGPU_SELECT_POS_SAMPLES")) // A/B knob
                v.samp_pos[b] = v.samp_pos[b] && std::atoi(e) != 0;
            // sentinel(s): the last block
            std::vector<uint32_t> tail(2, (uint32_t)(v.samp_pos[b] ? (v.nblocks - 1) * 7 + 6 : v.nblocks - 1));
            SG_CUDA(cudaMemcpyAsync(v.samp[b] + v.nsamp[b], tail.data(), 8, cudaMemcpyHostToDevice, s));
            if (m)
            {
                if (b)
                    bv_samples_kernel<1><<<blocks_for(v.nblocks), kThreads, 0, s>>>(abs_ones, blk_ones, nbits, v.nblocks, ls, v.samp[b], v.nsamp[b], v.blocks, v.samp_pos[b]);
                else
                    bv_samples_kernel<0><<<blocks_for(v.nblocks), kThreads, 0, s>>>(abs_ones, blk_ones, nbits, v.nblocks, ls, v.samp[b], v.nsamp[b], v.blocks, v.samp_pos[b]);
                SG_CUDA(cudaGetLastError());
            }
            SG_CUDA(cudaStreamSynchronize(s)); // `tail` must outlive the copy
        }
    }
    SG_CUDA(cudaStreamSynchronize(s));
    pool.release(blk_ones);
    pool.release(abs_ones);
    pool.release(tmp);

    if (flags & SDSLGPU_F_SDSL_LAYOUT)
    {
        v.words = words;
        // keep the caller's bits past nbits exactly as given, like the reference does
        for (int b = 0; b < 2; ++b)
            SG_TRY(bv_build_sdsl_rank_table(pool, v, b, s));
    }
    else
    {
        pool.release(words);
    }
    return SDSLGPU_OK;
}

// ------------------------------------------------------------------------------------------------
// two-bit patterns: rank_support_v<10|01|00|11, 2> / select_support_mcl<..., 2> (rank_support.hpp:161-284,
// select_support.hpp:204-405).  Every trait function of the reference is a popcount / select over the word of
// "the pattern ENDS at this bit" indicators (bits.hpp:565-583 map10 / map01, and the 00 / 11 analogues), with the
// carry = msb of the previous word and init_carry() for word 0.  So the supports of pattern p over v ARE the
// one-bit supports over that indicator vector: build it once from the sector blocks and reuse every kernel.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t bv_chunk32(bvblock const * __restrict__ blocks, uint64_t c)
{
    // 224 = 7 * 32: an aligned 32-bit chunk of the original vector never straddles a block
    return ld_nc_u32(&blocks[c / 7].d[c % 7]);
}

__global__ void bv_pattern_words_kernel(bvblock const * __restrict__ blocks, uint64_t nbits, int pat, uint64_t * __restrict__ words, uint64_t nwords)
{
    uint64_t const stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nwords; k += stride)
    {
        uint64_t nchunks = (nbits + 31) / 32;
        uint64_t lo = bv_chunk32(blocks, 2 * k), hi = (2 * k + 1 < nchunks) ? bv_chunk32(blocks, 2 * k + 1) : 0u;
        uint64_t x = lo | (hi << 32);
        uint64_t c = k ? (uint64_t)(bv_chunk32(blocks, 2 * k - 1) >> 31) : (uint64_t)(pat == SDSLGPU_PAT_01 || pat == SDSLGPU_PAT_00);
        uint64_t y = (x << 1) | c, m;
        switch (pat)
        {
        case SDSLGPU_PAT_10:
            m = y & ~x;
            break;
        case SDSLGPU_PAT_01:
            m = (x ^ y) & x;
            break;
        case SDSLGPU_PAT_00:
            m = ~(x | y);
            break;
        default:
            m = x & y;
            break;
        }
        uint64_t base = k * 64;
        if (base + 64 > nbits) // occurrences end inside the vector (util.hpp:689-726, select_support.hpp:286-301)
            m &= (nbits > base) ? lo_set64((uint32_t)(nbits - base)) : 0ull;
        words[k] = m;
    }
}

int bv_build_pattern(DevicePool & pool, BvImage const & src, int pat, BvImage & dst, cudaStream_t s)
{
    uint64_t nwords = (src.nbits + 63) >> 6;
    uint64_t * words = nullptr;
    SG_TRY(pool.alloc_t(&words, nwords + 2));
    if (nwords)
    {
        bv_pattern_words_kernel<<<blocks_for(nwords), kThreads, 0, s>>>(src.blocks, src.nbits, pat, words, nwords);
        SG_CUDA(cudaGetLastError());
    }
    int st = bv_build(pool, dst, SDSLGPU_F_DEFAULT, words, true, src.nbits, s);
    pool.release(words);
    return st;
}

int bv_build_sdsl_rank_table(DevicePool & pool, BvImage & v, int b, cudaStream_t s, bool v5)
{
    // rank_support_v.hpp:79,84: 2 words for an empty vector, else 2 * (((n+63)>>9) + 1); v5: >> 11 (rank_support_v5.hpp:73-79)
    uint64_t nsuper = (v.nbits == 0) ? 1 : ((v.nbits + 63) >> (v5 ? 11 : 9)) + 1;
    v.table_words = 2 * nsuper;
    SG_TRY(pool.alloc_t(&v.rank_table[b], v.table_words + 2));
    uint32_t * sb_total = nullptr;
    uint64_t * abs = nullptr;
    uint64_t * tmp = nullptr;
    SG_TRY(pool.alloc_t(&sb_total, nsuper));
    SG_TRY(pool.alloc_t(&abs, nsuper + 1));
    SG_TRY(pool.alloc_t(&tmp, scan_tmp_words(nsuper)));
    if (v5 && b)
        sdsl_table5_rel_kernel<1><<<blocks_for(nsuper), kThreads, 0, s>>>(v.words, v.nwords, nsuper, v.rank_table[b], sb_total);
    else if (v5)
        sdsl_table5_rel_kernel<0><<<blocks_for(nsuper), kThreads, 0, s>>>(v.words, v.nwords, nsuper, v.rank_table[b], sb_total);
    else if (b)
        sdsl_table_rel_kernel<1><<<blocks_for(nsuper), kThreads, 0, s>>>(v.words, v.nwords, nsuper, v.rank_table[b], sb_total);
    else
        sdsl_table_rel_kernel<0><<<blocks_for(nsuper), kThreads, 0, s>>>(v.words, v.nwords, nsuper, v.rank_table[b], sb_total);
    SG_CUDA(cudaGetLastError());
    SG_CUDA(exclusive_scan(sb_total, nsuper, abs, tmp, s));
    sdsl_table_abs_kernel<<<blocks_for(nsuper), kThreads, 0, s>>>(abs, nsuper, v.rank_table[b]);
    SG_CUDA(cudaGetLastError());
    SG_CUDA(cudaStreamSynchronize(s));
    pool.release(sb_total);
    pool.release(abs);
    pool.release(tmp);
    return SDSLGPU_OK;
}

int bv_rank_device(BvImage const & v, uint32_t flags, int b, uint64_t const * idx, uint64_t n, uint64_t * out, cudaStream_t s, Fan const * fan, bool * fanned)
{
    if (fanned)
        *fanned = false;
    if (n == 0)
        return SDSLGPU_OK;
    if (!(flags & SDSLGPU_F_SDSL_LAYOUT) && bv_binned_wanted(v, n))
    {
        bool done = false;
        SG_TRY(bv_rank_binned_device(v, b, idx, n, out, s, &done, fan));
        if (done)
        {
            if (fanned)
                *fanned = fan && fan->n;
            return SDSLGPU_OK;
        }
    }
    constexpr int ILP = 2;
    unsigned grid = grid_for(n, ILP);
    if (flags & SDSLGPU_F_SDSL_LAYOUT)
    {
        if (b)
            bv_rank_sdsl_kernel<1, ILP><<<grid, kThreads, 0, s>>>(v.words, v.rank_table[1], v.nbits, idx, n, out);
        else
            bv_rank_sdsl_kernel<0, ILP><<<grid, kThreads, 0, s>>>(v.words, v.rank_table[0], v.nbits, idx, n, out);
    }
    else if (fan && fan->n)
    {
        if (fan->width)
        {
            if (b)
                bv_rank_kernel<1, ILP, 2><<<grid, kThreads, 0, s>>>(v.blocks, v.top, v.nbits, idx, n, out, *fan);
            else
                bv_rank_kernel<0, ILP, 2><<<grid, kThreads, 0, s>>>(v.blocks, v.top, v.nbits, idx, n, out, *fan);
        }
        else if (b)
            bv_rank_kernel<1, ILP, 1><<<grid, kThreads, 0, s>>>(v.blocks, v.top, v.nbits, idx, n, out, *fan);
        else
            bv_rank_kernel<0, ILP, 1><<<grid, kThreads, 0, s>>>(v.blocks, v.top, v.nbits, idx, n, out, *fan);
        if (fanned)
            *fanned = true;
    }
    else
    {
        if (b)
            bv_rank_kernel<1, ILP, 0><<<grid, kThreads, 0, s>>>(v.blocks, v.top, v.nbits, idx, n, out, Fan{});
        else
            bv_rank_kernel<0, ILP, 0><<<grid, kThreads, 0, s>>>(v.blocks, v.top, v.nbits, idx, n, out, Fan{});
    }
    SG_CUDA(cudaGetLastError());
    return SDSLGPU_OK;
}

int bv_select_device(BvImage const & v, int b, uint64_t const * idx, uint64_t n, uint64_t * out, cudaStream_t s, Fan const * fan, bool * fanned)
{
    if (fanned)
        *fanned = false;
    if (n == 0)
        return SDSLGPU_OK;
    if (v.samp[b] == nullptr)
    {
        set_error("select requested on a handle created with SDSLGPU_F_NO_SELECT");
        return SDSLGPU_ENOTSUP;
    }
    if (bv_binned_wanted(v, n, true, b))
    {
        bool done = false;
        SG_TRY(bv_select_binned_device(v, b, idx, n, out, s, &done, fan));
        if (done)
        {
            if (fanned)
                *fanned = fan && fan->n;
            return SDSLGPU_OK;
        }
    }
    unsigned grid = grid_for(n);
    uint64_t args = b ? v.ones : v.nbits - v.ones;
    if (fan && fan->n)
    {
        if (fan->width)
        {
            if (b)
                bv_select_kernel<1, 2><<<grid, kThreads, 0, s>>>(bv_view(v), args, idx, n, out, *fan);
            else
                bv_select_kernel<0, 2><<<grid, kThreads, 0, s>>>(bv_view(v), args, idx, n, out, *fan);
        }
        else if (b)
            bv_select_kernel<1, 1><<<grid, kThreads, 0, s>>>(bv_view(v), args, idx, n, out, *fan);
        else
            bv_select_kernel<0, 1><<<grid, kThreads, 0, s>>>(bv_view(v), args, idx, n, out, *fan);
        if (fanned)
            *fanned = true;
    }
    else if (b)
        bv_select_kernel<1, 0><<<grid, kThreads, 0, s>>>(bv_view(v), args, idx, n, out, Fan{});
    else
        bv_select_kernel<0, 0><<<grid, kThreads, 0, s>>>(bv_view(v), args, idx, n, out, Fan{});
    SG_CUDA(cudaGetLastError());
    return SDSLGPU_OK;
}

// Builds the select sectors of pattern b of one of a handle's bit-vector images the first time a select batch large
// enough for the locality-ordered pipeline arrives (the handle is logically const; same pattern as the two-bit pattern
// images in api.cu).  Not built — and the sampled select keeps serving — for handles created with SDSLGPU_F_COMPACT,
// vectors beyond 2^36 bits, densities below ~5 %, or when the memory is not there.
int bv_ensure_select_sectors_image(sdslgpu_handle const * ch, BvImage const & cv, int b, uint64_t reserve_bytes)
{
    if ((b != 0 && b != 1) || cv.sect_tried[b] || cv.samp[b] == nullptr)
        return SDSLGPU_OK;
    sdslgpu_handle * h = const_cast<sdslgpu_handle *>(ch);
    std::lock_guard<std::mutex> lock(h->pat_mu);
    BvImage & v = const_cast<BvImage &>(cv); // one of h's own images
    if (v.sect_tried[b])
        return SDSLGPU_OK;
    v.sect_tried[b] = true;
    uint64_t const args = b ? v.ones : v.nbits - v.ones;
    uint32_t stride = bv_sect_stride(args, v.nbits);
    if (char const * e = std::getenv("SDSLGPU_SELECT_SECTOR_STRIDE")) // tuning knob
        stride = (uint32_t)std::atoi(e);
    bool off = (h->flags & SDSLGPU_F_COMPACT) != 0 || v.nbits > (1ull << 36) || stride < 2 || args == 0; // (the magic of 1 is 2^64)
    if (char const * e = std::getenv("SDSLGPU_SELECT_SECTORS")) // A/B knob
        off = off || std::atoi(e) == 0;
    if (off)
        return SDSLGPU_OK;
    uint64_t const nsect = (args - 1) / stride + 1;
    size_t free_b = 0, total_b = 0;
    DeviceGuard g(h->device);
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess || (nsect + 1) * sizeof(bvblock) + reserve_bytes + (1ull << 30) > free_b)
    {
        cudaGetLastError();
        return SDSLGPU_OK; // not enough room next to the batch's own scratch and 1 GiB for the caller: keep the sampled select
    }
    bvblock * sect = nullptr;
    if (h->pool.alloc_t(&sect, nsect + 1) != SDSLGPU_OK)
        return SDSLGPU_OK;
    if (b)
        bv_select_sectors_kernel<1><<<blocks_for(nsect), kThreads, 0, nullptr>>>(bv_view(v), v.nblocks, args, stride, nsect, sect);
    else
        bv_select_sectors_kernel<0><<<blocks_for(nsect), kThreads, 0, nullptr>>>(bv_view(v), v.nblocks, args, stride, nsect, sect);
    SG_CUDA(cudaGetLastError());
    SG_CUDA(cudaStreamSynchronize(nullptr));
    v.sect_stride[b] = stride;
    v.sect_magic[b] = bv_sect_magic(stride);
    v.nsect[b] = nsect;
    v.sect[b] = sect; // last: a concurrent reader either sees no sectors or complete ones
    return SDSLGPU_OK;
}

// KIND_BV handles: a no-op unless a select batch of n queries would run through the pipeline
int bv_ensure_select_sectors(sdslgpu_handle const * h, int b, uint64_t n)
{
    // (the batch sizes from which a one-gather select pays in the pipeline: what the sectors will make of this select)
    if ((b != 0 && b != 1) || h->bv.sect_tried[b] || !bin_wanted(h->bv.order, h->bv.nblocks * sizeof(bvblock), n, kBinRankDensity))
        return SDSLGPU_OK;
    return bv_ensure_select_sectors_image(h, h->bv, b, 16 * n); // 14 bytes of pipeline scratch per query (binned.cuh)
}

int bv_access_device(BvImage const & v, uint64_t const * idx, uint64_t n, uint64_t * out, cudaStream_t s)
{
    if (n == 0)
        return SDSLGPU_OK;
    bv_access_kernel<<<grid_for(n), kThreads, 0, s>>>(v.blocks, v.nbits, idx, n, out);
    SG_CUDA(cudaGetLastError());
    return SDSLGPU_OK;
}

} // namespace sdslgpu
