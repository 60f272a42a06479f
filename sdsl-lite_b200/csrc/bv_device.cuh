// bv_device.cuh — device-side view of a bit-vector image and the per-query primitives every kernel
// composes: rank1 at a position (one sector gather), bit access, and select by sampled hint + block scan.
// Used by bv.cu (plain rank/select), wt.cu (wavelet-tree levels), fm.cu (backward search, LF walks),
// sd.cu (Elias-Fano high part).
#pragma once
#include <cmath>

#include "common.cuh"

namespace sdslgpu
{

// plain-old-data view passed to kernels by value
struct BvView
{
    bvblock const * blocks;
    uint64_t const * top;
    uint32_t const * samp[2];
    uint32_t log_s[2];
    uint64_t nbits;
    uint64_t ones;
    uint32_t interp[2]; // 1: start the block search at the interpolated position between two samples
    uint32_t samp_pos[2]; // 1: samp[] holds (position of the sampled bit) >> 5 instead of the index of its block
    bvblock const * sect[2]; // select sectors (below) or nullptr
    uint32_t sect_stride[2]; // S: B-bits per sector
    uint64_t sect_magic[2];  // floor(2^64 / S) + 1: key / S = umul64hi(key, magic) for key < 2^64 / S
};

// ------------------------------------------------------------------------------------------------
// Select sectors: select in ONE 32-byte gather (large, dense vectors; built on the first large select batch).
// Sector j of pattern B serves the B-bits number j*S + 1 .. (j+1)*S, S = sect_stride:
//   cnt  : c0 = (position of B-bit number j*S + 1) >> 5, the 32-bit chunk it lies in; bit 31 = "does not fit"
//   d[7] : the 224 bits of the vector from chunk c0 on — complemented for B = 0, so the B-bits are the set bits —
//          with the bits before B-bit j*S + 1 cleared: the (r + 1)-th set bit of d is B-bit number j*S + r + 1.
// S is the largest stride whose S B-bits still fit the record at the vector's density with 3.5 standard deviations to
// spare (81 at density 1/2); where a sector's S B-bits need more than the 193 - 224 positions it holds (sparse
// stretches) bit 31 is set and the query goes the sampled way (bv_select).  The sampled select needs two dependent
// gathers per query (sample pair, block) and a third one for 5 % of them; the request port between an SM's L1 and the
// crossbar, not HBM, bounds it.  This structure trades memory for that: 32 bytes per S B-bits = 1.6 bits
// per bit of a half-dense vector (the reference's select_support_mcl: 0.2).
// ------------------------------------------------------------------------------------------------
static constexpr uint32_t kSectOverflow = 0x80000000u;

// B-bits per sector for a vector of `args` B-bits among `nbits`: the span of S B-bits at density d has mean S / d and
// variance S (1 - d) / d^2; the largest S with mean + 3.5 sigma <= 208 (the record holds 193 - 224 positions, depending
// on where in its first chunk the first B-bit lies).  Below 8 (density under ~8 %) a sector is no better than a
// position list: 0 is returned and no sectors are built.
inline uint32_t bv_sect_stride(uint64_t args, uint64_t nbits)
{
    if (args == 0 || nbits == 0)
        return 0;
    double const d = (double)args / (double)nbits, a = 3.5 * std::sqrt(1.0 - (d < 1.0 ? d : 1.0));
    double const x = (-a + std::sqrt(a * a + 4.0 * 208.0 * d)) / 2.0;
    uint32_t const S = (uint32_t)(x * x);
    return S >= 8 ? (S > 208 ? 208u : S) : 0u;
}
inline uint64_t bv_sect_magic(uint32_t stride)
{
    return ~0ull / stride + 1; // floor(2^64 / S) + 1 (also when S divides 2^64); S >= 2
}

// number of 1-bits in [0, pos), 0 <= pos <= nbits: one 32-byte sector gather
// (device form of rank_support_v<1>::rank, rank_support_v.hpp:129-139)
__device__ __forceinline__ uint64_t bv_rank1(BvView const & v, uint64_t pos)
{
    uint64_t blk = pos / kBlockBits;
    uint32_t rem = (uint32_t)(pos - blk * kBlockBits);
    uint32_t cnt, d[7];
    ld_block_half_line(v.blocks + blk, cnt, d);
    return __ldg(v.top + (blk >> kSuperShift)) + cnt + block_prefix_popc(d, rem);
}

// rank1(pos) and the bit AT pos from the same sector (pos < nbits): what wt access / LF need per level
__device__ __forceinline__ uint64_t bv_rank1_and_bit(BvView const & v, uint64_t pos, uint32_t & bit)
{
    uint64_t blk = pos / kBlockBits;
    uint32_t rem = (uint32_t)(pos - blk * kBlockBits);
    uint32_t cnt, d[7];
    ld_block(v.blocks + blk, cnt, d);
    uint32_t w = 0;
#pragma unroll
    for (uint32_t j = 0; j < 7; ++j)
        w = (j == (rem >> 5)) ? d[j] : w;
    bit = (w >> (rem & 31u)) & 1u;
    return __ldg(v.top + (blk >> kSuperShift)) + cnt + block_prefix_popc(d, rem);
}

__device__ __forceinline__ uint32_t bv_bit(BvView const & v, uint64_t pos)
{
    uint64_t blk = pos / kBlockBits;
    uint32_t rem = (uint32_t)(pos - blk * kBlockBits);
    return (ld_nc_u32(&v.blocks[blk].d[rem >> 5]) >> (rem & 31u)) & 1u;
}

template <int B>
__device__ __forceinline__ uint64_t abs_before(bvblock const * __restrict__ blocks, uint64_t const * __restrict__ top, uint64_t b)
{
    uint64_t a1 = __ldg(top + (b >> kSuperShift)) + __ldg(&blocks[b].cnt);
    return B ? a1 : b * kBlockBits - a1;
}

// position of the i-th (1-based) B-bit, given that it lies in a block of [lo, hi] and fewer than i B-bits precede
// block lo.  r / 2^log_span = how far i lies between the two bracketing hints in B-bit count; with `interp` the
// search starts at the interpolated block — for the sparse hint tables that stay on chip (spans up to 2^14 B-bits)
// this lands in the right sector block 75-90 % of the time on random data; any miss is repaired by walking /
// bisecting on the block counts, so the result never depends on the guess.
template <int B>
__device__ __forceinline__ uint64_t bv_select_from(BvView const & v, uint64_t i, uint64_t lo, uint64_t hi, uint64_t g);
// the same search, additionally handing out the block it ended in: g_out = its index, d_out = its 224 payload bits
// (callers that look at the neighbourhood of the answer — sd_vector rank — read them from registers, not from memory)
template <int B>
__device__ __forceinline__ uint64_t bv_select_from(BvView const & v, uint64_t i, uint64_t lo, uint64_t hi, uint64_t g, uint32_t (&d_out)[7], uint64_t & g_out);

template <int B>
__device__ __forceinline__ uint64_t bv_select_between(BvView const & v, uint64_t i, uint64_t lo, uint64_t hi, uint64_t r, uint32_t log_span, bool interp)
{
    return bv_select_from<B>(v, i, lo, hi, interp ? lo + (((hi - lo) * r + (1ull << log_span >> 1)) >> log_span) : lo);
}

template <int B>
__device__ __forceinline__ uint64_t bv_select_from(BvView const & v, uint64_t i, uint64_t lo, uint64_t hi, uint64_t g)
{
    uint32_t d[7];
    uint64_t g_out;
    return bv_select_from<B>(v, i, lo, hi, g, d, g_out);
}

// the search proper: first probe at block g in [lo, hi], then walk / bisect
template <int B>
__device__ __forceinline__ uint64_t bv_select_from(BvView const & v, uint64_t i, uint64_t lo, uint64_t hi, uint64_t g, uint32_t (&d)[7], uint64_t & g_out)
{
    bvblock const * __restrict__ blocks = v.blocks;
    uint64_t const * __restrict__ top = v.top;
    uint32_t cnt;
    ld_block(blocks + g, cnt, d);
    uint64_t a1 = __ldg(top + (g >> kSuperShift)) + cnt;
    uint64_t before = B ? a1 : g * kBlockBits - a1;
    uint32_t steps = 0;
    if (before >= i)
    { // overshoot: the answer is in [lo, g-1].  Walk left; after three near misses with a long way to go (clustered
      // data, wide hint spans) bisect on the block counts, then finish the walk
        for (;;)
        {
            --g;
            if (++steps == 4 && g - lo > 3)
            {
                uint64_t h = g; // before(h + 1) >= i, before(lo) < i
                while (h - lo > 3)
                {
                    uint64_t mid = (lo + h + 1) >> 1;
                    if (abs_before<B>(blocks, top, mid) < i)
                        lo = mid;
                    else
                        h = mid - 1;
                }
                g = h;
            }
            ld_block(blocks + g, cnt, d);
            a1 = __ldg(top + (g >> kSuperShift)) + cnt;
            before = B ? a1 : g * kBlockBits - a1;
            if (before < i)
                break; // the first block from the right whose prefix count drops below i holds the answer
        }
    }
    uint64_t need = i - before;
    uint32_t c = block_popc<B>(d);
    while (need > c)
    { // undershoot: walk right, same escape to a bisection
        need -= c;
        ++g;
        if (++steps == 4 && hi - g > 8)
        {
            uint64_t l = g, h = hi; // before(l) < i
            while (h - l > 3)
            {
                uint64_t mid = (l + h + 1) >> 1;
                if (abs_before<B>(blocks, top, mid) < i)
                    l = mid;
                else
                    h = mid - 1;
            }
            g = l;
            ld_block(blocks + g, cnt, d);
            a1 = __ldg(top + (g >> kSuperShift)) + cnt;
            before = B ? a1 : g * kBlockBits - a1;
            need = i - before;
        }
        else
            ld_block(blocks + g, cnt, d);
        c = block_popc<B>(d);
    }
    g_out = g;
    return g * kBlockBits + block_select<B>(d, (uint32_t)need);
}


template <int B>
__device__ __forceinline__ uint64_t bv_select(BvView const & v, uint64_t i, uint32_t (&d_out)[7], uint64_t & g_out);

// position of the i-th (1-based) B-bit, given 1 <= i <= #B-bits
// (device form of select_support_mcl<B>::select, select_support_mcl.hpp:384-439: sampled hint, then a
//  scan over block counts instead of the reference's word scan)
template <int B>
__device__ __forceinline__ uint64_t bv_select(BvView const & v, uint64_t i)
{
    uint32_t d[7];
    uint64_t g;
    return bv_select<B>(v, i, d, g);
}

template <int B>
__device__ __forceinline__ uint64_t bv_select(BvView const & v, uint64_t i, uint32_t (&d_out)[7], uint64_t & g_out)
{
    uint32_t const * __restrict__ samp = v.samp[B];
    uint32_t const log_s = v.log_s[B];
    uint64_t j = (i - 1) >> log_s;
    uint2 s2;
    // samp[j], samp[j+1]: one 8-byte load when j is even
    uint64_t lo, hi;
    if ((j & 1) == 0)
    {
        s2 = __ldg(reinterpret_cast<uint2 const *>(samp + j));
        lo = s2.x;
        hi = s2.y;
    }
    else
    {
        lo = __ldg(samp + j);
        hi = __ldg(samp + j + 1);
    }
    uint64_t const r = (i - 1) & ((1ull << log_s) - 1);
    if (v.samp_pos[B])
    { // position-valued samples: lo / hi are 32-bit chunk indices (224 = 7 * 32: a chunk never straddles two blocks), so
      // the interpolation is not quantised to whole blocks at either end — first-probe misses 18 % -> 5 % on random data.
      // Everything stays in 32-bit registers: chunk indices are < 2^32 and r < 2^log_s <= 2^16, so one 32 x 32 -> 64
      // multiply, and the three divisions by 7 are 32-bit multiply-high sequences instead of 64-bit ones.
        uint32_t const lo32 = (uint32_t)lo, hi32 = (uint32_t)hi;
        uint32_t const p = lo32 + (uint32_t)(((uint64_t)(hi32 - lo32) * (uint32_t)r + (1ull << log_s >> 1)) >> log_s);
        return bv_select_from<B>(v, i, lo32 / 7u, hi32 / 7u, p / 7u, d_out, g_out);
    }
    return bv_select_from<B>(v, i, lo, hi, v.interp[B] ? lo + (((hi - lo) * r + (1ull << log_s >> 1)) >> log_s) : lo, d_out, g_out);
}

// B-bit number key + 1 from its select sector; *fits = false (and nothing else done) when the sector is marked
template <int B>
__device__ __forceinline__ uint64_t bv_select_sector(BvView const & v, uint64_t key, bool & fits)
{
    uint64_t const j = __umul64hi(key, v.sect_magic[B]); // key / S
    uint32_t c0, d[7];
    ld_block(v.sect[B] + j, c0, d);
    fits = (c0 & kSectOverflow) == 0;
    if (!fits)
        return 0;
    return ((uint64_t)c0 << 5) + block_select<1>(d, (uint32_t)(key - j * v.sect_stride[B]) + 1u);
}

// select by whichever structure the image has: its select sector if there is one and the query fits it, else samples
template <int B>
__device__ __forceinline__ uint64_t bv_select_any(BvView const & v, uint64_t i)
{
    if (v.sect[B])
    {
        bool fits;
        uint64_t const r = bv_select_sector<B>(v, i - 1, fits);
        if (fits)
            return r;
    }
    return bv_select<B>(v, i);
}

// what the build kernel stores for sector j (bv.cu bv_select_sectors_kernel; the CPU tests build their images with it)
template <int B>
__device__ __forceinline__ void bv_make_sector(BvView const & v, uint64_t nblocks, uint64_t args, uint32_t stride, uint64_t j, uint32_t & c0_out, uint32_t (&d)[7])
{
    uint64_t const pos0 = bv_select<B>(v, j * stride + 1);
    uint64_t const c0 = pos0 >> 5;
#pragma unroll
    for (uint32_t t = 0; t < 7; ++t)
    {
        uint64_t const c = c0 + t, blk = c / 7;
        uint32_t x = 0;
        if (blk < nblocks && c * 32 < v.nbits)
        {
            x = ld_nc_u32(&v.blocks[blk].d[c - blk * 7]);
            if (!B)
            {
                x = ~x;
                if (c * 32 + 32 > v.nbits) // the vector ends inside this chunk: what lies behind it is not a 0-bit
                    x &= (1u << (uint32_t)(v.nbits - c * 32)) - 1u;
            }
        }
        d[t] = x;
    }
    d[0] &= ~((1u << ((uint32_t)pos0 & 31u)) - 1u);
    uint64_t const left = args - j * stride;
    uint32_t const need = (uint32_t)(left < stride ? left : stride);
    c0_out = (uint32_t)c0 | (block_popc<1>(d) < need ? kSectOverflow : 0u);
}

} // namespace sdslgpu
