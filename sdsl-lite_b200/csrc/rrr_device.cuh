// rrr_device.cuh — device-side view of an rrr_vector<63> image and its per-query primitives (rank, bit, select),
// shared by rrr.cu (plain rank/select/access batches) and the wavelet-tree / FM-index kernels when the tree's
// bit vector is stored H0-compressed (wt_huff<rrr_vector<63>>).  Layout: see rrr.cu.
#pragma once
#include <cstring>
#include <vector>

#include "common.cuh"

namespace sdslgpu
{

static constexpr uint32_t kBs = 63; // t_bs
static constexpr uint32_t kK = 32;  // t_k
static constexpr uint64_t kInvBit = 1ull << 63;
static constexpr uint32_t kRecWords = 8;
// select hints: one per 2^hint_shift b-bits, the stride chosen per vector and bit value so that two neighbouring hints
// bracket about four superblocks whatever the density (rrr_hint_shift; a fixed stride of 2^13 made a 1 %-dense vector
// bisect over ~400 records, nine dependent gathers)
__host__ __device__ __forceinline__ uint32_t rrr_hint_shift(uint64_t args, uint64_t nsuper)
{
    uint32_t s = 0;
    while (s < 20 && (args >> s) > nsuper / 4 + 1)
        ++s;
    return s;
}

// C(n, k) for n <= 62, 0 for k > n (rrr_helper.hpp:193-237), split by magnitude: every C(n, k) with n <= 33 is below
// 2^32, so the last 34 steps of a block's enumerative decode run in 32-bit registers on a table of half the size.
// 23.6 KB of shared memory per CTA (the full 64 x 64 x u64 table was 32 KB).
#ifndef RRR_SPARSE_PATH
#define RRR_SPARSE_PATH 1 // 0: every block is decoded position by position (A/B partner, tools/variants.sh)
#endif
static constexpr uint32_t kBinomSplit = 34;
struct RrrTables
{
    uint64_t hi[63 - kBinomSplit][64]; // C(n, k), n = 34 .. 62
    uint32_t lo[kBinomSplit][64];      // C(n, k), n = 0 .. 33
    uint8_t space[64];                 // bits of an offset of class k: 0 if C(63,k) == 1 else hi(C(63,k)) + 1 (:286-293)
};
static_assert(sizeof(RrrTables) % 16 == 0, "staged with 16-byte copies");

__host__ __device__ __forceinline__ uint64_t rrr_binom(RrrTables const * t, uint32_t n, uint32_t k)
{
    return n >= kBinomSplit ? t->hi[n - kBinomSplit][k] : (uint64_t)t->lo[n][k];
}

#ifndef SDSLGPU_HOST_EMU
__device__ __forceinline__ void stage_rrr(RrrTables const * __restrict__ g, RrrTables * s)
{
    uint4 const * src = reinterpret_cast<uint4 const *>(g);
    uint4 * dst = reinterpret_cast<uint4 *>(s);
    for (uint32_t k = threadIdx.x; k < sizeof(RrrTables) / 16; k += blockDim.x)
        dst[k] = __ldg(src + k);
    __syncthreads();
}
#endif

// Enumerative decode of one block (k ones among 63 positions, offset nr), the inverse of bin_to_nr
// (rrr_helper.hpp:346-366): position p holds a one iff nr >= C(62 - p, k) for the k ones still to place; then
// nr -= C(62 - p, k), --k.  Before step p, nr < C(63 - p, k) — so from p = 29 on nr < C(34, k) < 2^32 and the walk
// continues in 32-bit arithmetic.  The reference's decode_popcount / decode_bit / decode_select (rrr_helper.hpp:369-649)
// are all functions of this walk; nothing here materialises the 63-bit word.
static constexpr uint32_t kWide = 63 - kBinomSplit; // steps p = 0 .. 28 compare 64-bit binomials

// One step of the walk at position p, k ones left: is there a one?  (k == 0 needs no special case: C(n, 0) = 1 > nr = 0.)
// Written on byte offsets into the two tables so that the unrolled loops below keep only `nr` and the column offset
// live: LDS, compare, conditional subtract, conditional column step — 7 instructions per 64-bit step, 5 per 32-bit step.
#ifdef SDSLGPU_HOST_EMU
#define SG_UNROLL4
// if nr >= c: nr -= c, --k; returns whether a one was placed
inline bool rrr_step64(uint64_t & nr, uint32_t & k, uint64_t c)
{
    bool const one = nr >= c;
    nr -= one ? c : 0ull;
    k -= one ? 1u : 0u;
    return one;
}
inline bool rrr_step32(uint32_t & r, uint32_t & k, uint32_t c)
{
    bool const one = r >= c;
    r -= one ? c : 0u;
    k -= one ? 1u : 0u;
    return one;
}
#else
#define SG_UNROLL4 _Pragma("unroll 4")
// predicated forms: one compare, then the subtraction and the decrement under its predicate (the compiler's own
// translation of the ternaries selects zero-or-c into temporaries first: 4 more instructions per 64-bit step)
__device__ __forceinline__ bool rrr_step64(uint64_t & nr, uint32_t & k, uint64_t c)
{
    uint32_t one;
    asm("{\n .reg .pred p;\n setp.ge.u64 p, %0, %3;\n @p sub.u64 %0, %0, %3;\n @p sub.u32 %1, %1, 1;\n selp.u32 %2, 1, 0, p;\n}"
        : "+l"(nr), "+r"(k), "=r"(one)
        : "l"(c));
    return one != 0;
}
__device__ __forceinline__ bool rrr_step32(uint32_t & r, uint32_t & k, uint32_t c)
{
    uint32_t one;
    asm("{\n .reg .pred p;\n setp.ge.u32 p, %0, %3;\n @p sub.u32 %0, %0, %3;\n @p sub.u32 %1, %1, 1;\n selp.u32 %2, 1, 0, p;\n}"
        : "+r"(r), "+r"(k), "=r"(one)
        : "r"(c));
    return one != 0;
}
#endif

// ------------------------------------------------------------------------------------------------
// Blocks with few ones (or few zeros): the ones are found one at a time instead of position by position.  With k ones
// left and every position p < p0 decided, the next one sits at the first p >= p0 with nr >= C(62 - p, k): C(n, k) grows
// with n, so that is n* = the largest n <= 62 - p0 with C(n, k) <= nr — a bisection over a column of the table (six
// probes at most) per ONE, where the walk pays one step per POSITION (31 on average, and a warp waits for its longest).
// At 1 - 5 % density (0.6 - 3 ones per block) that is the difference between ~100 and ~450 instructions per warp trip.
// A block with few zeros is its complement's mirror image: the complement of a block of class k and offset nr has
// class 63 - k and offset C(63, k) - 1 - nr (the enumeration is lexicographic, complementing reverses it).
// The search costs ~50 instructions per one, so it only wins for a handful of ones — and a warp that holds one lane on
// the walk pays for the walk anyway: the search is taken only when EVERY lane of the warp that arrives here together has
// such a block (measured, profiles/r02z_rrr_sparse_variants.jsonl: a per-lane switch at k <= 10 made 5 - 25 % density up
// to 60 % slower; with the warp-wide switch 1 % density gains 19 % in rank, 10 - 14 % in select, the rest is unchanged).
// Either path is correct for any class, so the answer never depends on the choice.  Vectors between 1/32 and 31/32 dense do
// not even ask (RrrView::try_sparse): the vote alone cost them 1 - 2 %.
// ------------------------------------------------------------------------------------------------
static constexpr uint32_t kSparseK = 5;
#ifdef SDSLGPU_HOST_EMU
inline bool rrr_warp_all(bool x)
{
    return x;
}
#else
__device__ __forceinline__ bool rrr_warp_all(bool x)
{
    return __all_sync(__activemask(), x) != 0;
}
#endif

// largest n in [lo, hi] with C(n, k) <= nr; requires C(lo, k) <= nr
__device__ __forceinline__ uint32_t rrr_largest_n(RrrTables const * t, uint32_t k, uint64_t nr, uint32_t lo, uint32_t hi)
{
    while (lo < hi)
    {
        uint32_t const mid = (lo + hi + 1) >> 1;
        if (rrr_binom(t, mid, k) <= nr)
            lo = mid;
        else
            hi = mid - 1;
    }
    return lo;
}

// class and offset of the complemented block
__device__ __forceinline__ void rrr_complement(RrrTables const * t, uint32_t & k, uint64_t & nr)
{
    uint64_t const c63 = rrr_binom(t, 62, k) + rrr_binom(t, 62, k - 1); // C(63, k), 1 <= k <= 62
    nr = c63 - 1 - nr;
    k = kBs - k;
}

// ones among positions [0, off) of a block with 1 <= k <= kSparseK ones (and the bit at `off`, off <= 62)
__device__ __forceinline__ uint32_t rrr_prefix_ones_sparse(RrrTables const * t, uint32_t k, uint64_t nr, uint32_t off, bool want_bit, uint32_t & bit)
{
    uint32_t cnt = 0, n_hi = 62;
    uint32_t const lim = 63 - off; // a one at position p < off has n = 62 - p >= lim
    while (k)
    {
        if (lim > n_hi || rrr_binom(t, lim, k) > nr)
            break; // the next one lies at or behind `off`
        uint32_t const n = rrr_largest_n(t, k, nr, lim, n_hi);
        nr -= rrr_binom(t, n, k);
        --k;
        ++cnt;
        if (n == 0)
            break;
        n_hi = n - 1;
    }
    bit = 0;
    if (want_bit && k) // the next one sits exactly at `off` iff C(62 - off, k) <= nr (62 - off <= n_hi holds here)
        bit = rrr_binom(t, 62 - off, k) <= nr;
    return cnt;
}

// position of the target-th (1-based) B-bit of a block with 1 <= k <= kSparseK ONES; requires that it exists
template <int B>
__device__ __forceinline__ uint32_t rrr_select_sparse(RrrTables const * t, uint32_t k, uint64_t nr, uint32_t target)
{
    uint32_t j = 0, n_hi = 62; // j = ones found so far
    while (k)
    {
        uint32_t const n = rrr_largest_n(t, k, nr, k - 1, n_hi); // C(k - 1, k) = 0 <= nr
        uint32_t const p = 62 - n;
        if (B)
        {
            if (++j == target)
                return p;
        }
        else
        {
            if (p - j >= target) // zeros before this one
                return target - 1 + j;
            ++j;
        }
        nr -= rrr_binom(t, n, k);
        --k;
        if (n == 0)
            break;
        n_hi = n - 1;
    }
    return target - 1 + j; // B = 0: behind the last one
}

// ones among positions [0, off) and, if want_bit, the bit at position off (off < 63 then)
__device__ __forceinline__ uint32_t rrr_prefix_ones(RrrTables const * t, uint32_t k, uint64_t nr, uint32_t off, bool want_bit, uint32_t & bit, bool try_sparse = true)
{
    if (k == 1)
    { // one one: it sits at position 62 - nr (nr = C(62 - p, 1)); a third of all blocks at 1 % density
        uint32_t const one_at = 62u - (uint32_t)nr;
        bit = want_bit && one_at == off;
        return one_at < off;
    }
#if RRR_SPARSE_PATH
    if (try_sparse && rrr_warp_all(k <= kSparseK || k >= kBs - kSparseK))
    {
        if (k <= kSparseK)
            return rrr_prefix_ones_sparse(t, k, nr, off, want_bit, bit);
        // few zeros: count them in the complement
        rrr_complement(t, k, nr);
        uint32_t const zeros = rrr_prefix_ones_sparse(t, k, nr, off, want_bit, bit);
        bit = want_bit ? 1u - bit : 0u;
        return off - zeros;
    }
#endif
    uint32_t const k0 = k;
    uint32_t const wide = off < kWide ? off : kWide; // steps [0, wide) on 64-bit binomials, [wide, off) on 32-bit ones
    {
        uint64_t const * row = &t->hi[kWide - 1][0]; // C(62 - p, .) for p = 0; one row back per step
        SG_UNROLL4
        for (uint32_t p = 0; p < wide; ++p, row -= 64)
            rrr_step64(nr, k, row[k]);
    }
    uint32_t r = (uint32_t)nr; // from step 29 on nr < C(34, k) < 2^32
    if (off > kWide)
    {
        uint32_t const * row = &t->lo[62 - kWide][0];
        SG_UNROLL4
        for (uint32_t p = kWide; p < off; ++p, row -= 64)
            rrr_step32(r, k, row[k]);
    }
    bit = 0;
    if (want_bit)
        bit = off < kWide ? (nr >= t->hi[kWide - 1 - off][k]) : (r >= t->lo[62 - off][k]);
    return k0 - k;
}

// position (0..62) of the target-th (1-based) B-bit of the block; requires target <= number of B-bits
template <int B>
__device__ __forceinline__ uint32_t rrr_select_in_block(RrrTables const * t, uint32_t k, uint64_t nr, uint32_t target, bool try_sparse = true)
{
    if (k == 1)
    {
        uint32_t const one_at = 62u - (uint32_t)nr;
        return B ? one_at : (target - 1 < one_at ? target - 1 : target);
    }
#if RRR_SPARSE_PATH
    if (try_sparse && rrr_warp_all(k <= kSparseK || k >= kBs - kSparseK))
    {
        if (k <= kSparseK)
            return rrr_select_sparse<B>(t, k, nr, target);
        // few zeros: the B-bits of the block are the (1 - B)-bits of its complement
        rrr_complement(t, k, nr);
        return rrr_select_sparse<1 - B>(t, k, nr, target);
    }
#endif
    {
        uint64_t const * row = &t->hi[kWide - 1][0];
        SG_UNROLL4
        for (uint32_t p = 0; p < kWide; ++p, row -= 64)
        {
            bool const one = rrr_step64(nr, k, row[k]);
            target -= (one == (B != 0)) ? 1u : 0u;
            if (target == 0)
                return p;
        }
    }
    uint32_t r = (uint32_t)nr;
    uint32_t const * row = &t->lo[62 - kWide][0];
    SG_UNROLL4
    for (uint32_t p = kWide; p < kBs; ++p, row -= 64)
    {
        bool const one = rrr_step32(r, k, row[k]);
        target -= (one == (B != 0)) ? 1u : 0u;
        if (target == 0)
            return p;
    }
    return kBs - 1; // not reached for a valid target
}

struct RrrView
{
    uint64_t size;
    uint64_t nblocks; // m_bt.size()
    uint64_t nsuper;  // m_btnrp.size(); `records` has nsuper + 1 entries (the last holds the totals)
    uint64_t ones;
    uint64_t const * btnr;    // packed offsets (m_btnr)
    uint64_t const * records; // 8 words per superblock, see the file header
    RrrTables const * tables;
    uint32_t const * hint[2]; // hint[b][j] = superblock holding the (j * 2^hint_shift[b] + 1)-th b-bit (+ sentinels)
    uint32_t hint_shift[2];
    uint32_t try_sparse; // the vector is sparse or dense enough (<= 1/32 ones or zeros) for whole warps to meet few-one blocks
};

struct RrrRecord
{
    uint64_t w[kRecWords];
};

#ifdef SDSLGPU_HOST_EMU
inline void ld_record(uint64_t const * records, uint64_t g, RrrRecord & r)
{
    std::memcpy(r.w, records + g * kRecWords, sizeof(r.w));
}
#else
__device__ __forceinline__ void ld_record(uint64_t const * __restrict__ records, uint64_t g, RrrRecord & r)
{
    uint64_t const * p = records + g * kRecWords;
    asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(r.w[0]), "=l"(r.w[1]), "=l"(r.w[2]), "=l"(r.w[3]) : "l"(p));
    asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(r.w[4]), "=l"(r.w[5]), "=l"(r.w[6]), "=l"(r.w[7]) : "l"(p + 4));
}
#endif

// stored class of block j (0..31) of a record
__host__ __device__ __forceinline__ uint32_t rec_class(uint64_t w2, uint64_t w3, uint64_t w4, uint32_t j)
{
    uint32_t bit = j * 6;
    if (bit < 60)
        return (uint32_t)(w2 >> bit) & 63u;
    if (bit == 60)
        return (uint32_t)((w2 >> 60) | (w3 << 4)) & 63u;
    if (bit < 124)
        return (uint32_t)(w3 >> (bit - 64)) & 63u;
    if (bit == 126)
        return (uint32_t)((w3 >> 62) | (w4 << 2)) & 63u;
    return (uint32_t)(w4 >> (bit - 128)) & 63u;
}

// prefix sums over whole quarters (8 blocks each): q in 0..3
__host__ __device__ __forceinline__ uint32_t rec_qones(uint64_t w5, uint32_t q)
{
    return q == 0 ? 0u : q == 1 ? (uint32_t)(w5 & 1023) : q == 2 ? (uint32_t)((w5 >> 10) & 1023) : (uint32_t)((w5 >> 20) & 2047);
}
__host__ __device__ __forceinline__ uint32_t rec_qbits(uint64_t w5, uint32_t q)
{
    return q == 0 ? 0u : q == 1 ? (uint32_t)((w5 >> 31) & 1023) : q == 2 ? (uint32_t)((w5 >> 41) & 1023) : (uint32_t)((w5 >> 51) & 2047);
}

// the eight stored classes of quarter q (blocks 8q .. 8q+7): 48 bits starting at bit 48 q of w2..w4
__host__ __device__ __forceinline__ uint64_t rec_quarter(uint64_t w2, uint64_t w3, uint64_t w4, uint32_t q)
{
    uint64_t const x = q == 0 ? w2 : q == 1 ? ((w2 >> 48) | (w3 << 16)) : q == 2 ? ((w3 >> 32) | (w4 << 32)) : (w4 >> 16);
    return x & 0xFFFFFFFFFFFFull;
}

// ones and offset bits of the first nblk (0..31) blocks of a superblock; also returns the class of block nblk itself
// (real, i.e. inversion undone).  Quarter prefixes come from w5; inside the quarter the stored classes are summed in
// one SWAR step (six-bit fields -> four 12-bit lanes -> one multiply) and the code lengths by at most 7 table reads,
// without a branch per class.
__device__ __forceinline__ uint32_t rec_prefix(RrrRecord const & r, RrrTables const * t, uint32_t nblk, bool inv, uint64_t & ones, uint64_t & p)
{
    uint32_t const q = nblk >> 3, rem = nblk & 7u;
    ones += rec_qones(r.w[5], q);
    p += rec_qbits(r.w[5], q);
    uint64_t const x = rec_quarter(r.w[2], r.w[3], r.w[4], q);
    uint64_t const low = x & ((1ull << (6u * rem)) - 1ull); // the classes of the rem blocks before block nblk
    uint64_t const kLanes = 0x03F03F03F03Full;               // fields 0, 2, 4, 6
    uint64_t const pair = (low & kLanes) + ((low >> 6) & kLanes);
    uint32_t const stored = (uint32_t)((pair * 0x001001001001ull) >> 36) & 0xFFFu;
    ones += inv ? (uint64_t)kBs * rem - stored : stored;
    uint32_t bits = 0;
#pragma unroll
    for (uint32_t j = 0; j < 7; ++j)
        bits += j < rem ? t->space[(uint32_t)(x >> (6u * j)) & 63u] : 0u; // space is symmetric: stored or real class (rrr_vector.hpp:528-541)
    p += bits;
    uint32_t const c = (uint32_t)(x >> (6u * rem)) & 63u;
    return inv ? kBs - c : c;
}

// ------------------------------------------------------------------------------------------------
// queries
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t rrr_rank1_one(RrrView const & v, RrrTables const * t, uint64_t i)
{
    uint64_t blk = i / kBs, g = blk / kK;
    uint32_t off = (uint32_t)(i - blk * kBs);
    RrrRecord r;
    ld_record(v.records, g, r);
    uint64_t d = r.w[6];
    if (d == 0)
        return r.w[0]; // uniform superblocks (rrr_vector.hpp:514-523); same result as the general path
    if (d == (uint64_t)kBs * kK)
        return r.w[0] + i - g * kK * kBs;
    bool inv = (r.w[1] & kInvBit) != 0;
    uint64_t p = r.w[1] & ~kInvBit, ones = r.w[0];
    uint32_t nblk = (uint32_t)(blk - g * kK);
    uint32_t const k = rec_prefix(r, t, nblk, inv, ones, p);
    if (off == 0 || k == 0)
        return ones;
    if (k == kBs)
        return ones + off;
    uint32_t sp = t->space[k], bit;
    return ones + rrr_prefix_ones(t, k, sp ? read_int(v.btnr, p, sp) : 0, off, false, bit, v.try_sparse != 0);
}

// rank1(pos) and the bit at pos (pos < size) from one record + one offset read
__device__ __forceinline__ uint64_t rrr_rank1_and_bit(RrrView const & v, RrrTables const * t, uint64_t i, uint32_t & bit)
{
    uint64_t blk = i / kBs, g = blk / kK;
    uint32_t off = (uint32_t)(i - blk * kBs);
    RrrRecord r;
    ld_record(v.records, g, r);
    bool inv = (r.w[1] & kInvBit) != 0;
    uint64_t p = r.w[1] & ~kInvBit, ones = r.w[0];
    uint32_t nblk = (uint32_t)(blk - g * kK);
    uint32_t const k = rec_prefix(r, t, nblk, inv, ones, p);
    if (k == 0 || k == kBs)
    {
        bit = k != 0;
        return ones + (k ? off : 0u);
    }
    uint32_t sp = t->space[k];
    return ones + rrr_prefix_ones(t, k, sp ? read_int(v.btnr, p, sp) : 0, off, true, bit, v.try_sparse != 0);
}

// position of the i-th (1-based) B-bit, 1 <= i <= #B-bits (rrr_vector.hpp:639-726)
template <int B>
__device__ __forceinline__ uint64_t rrr_select_one(RrrView const & v, RrrTables const * t, uint64_t i)
{
    // superblock g with count_before(g) < i <= count_before(g + 1)   (:643-655), bracketed by the hints
    uint64_t hj = (i - 1) >> v.hint_shift[B];
    uint64_t begin = __ldg(v.hint[B] + hj), end = (uint64_t)__ldg(v.hint[B] + hj + 1) + 1;
    while (end - begin > 1)
    {
        uint64_t mid = (begin + end) >> 1;
        uint64_t rk = __ldg(v.records + mid * kRecWords);
        uint64_t c = B ? rk : mid * kBs * kK - rk;
        if (c >= i)
            end = mid;
        else
            begin = mid;
    }
    RrrRecord r;
    ld_record(v.records, begin, r);
    uint64_t cnt = B ? r.w[0] : begin * kBs * kK - r.w[0];
    uint64_t d = r.w[6];
    if (B ? (d == (uint64_t)kBs * kK) : (d == 0))
        return begin * kK * kBs + (i - cnt - 1); // all-ones / all-zeros superblock (:658-663, :703-706)
    bool inv = (r.w[1] & kInvBit) != 0;
    uint64_t p = r.w[1] & ~kInvBit;
    // skip whole quarters, then scan at most 8 classes
    uint32_t quarter = 0;
#pragma unroll
    for (uint32_t qq = 1; qq < 4; ++qq)
    {
        uint32_t o = rec_qones(r.w[5], qq);
        uint64_t c = B ? o : qq * 8 * kBs - o;
        if (cnt + c < i)
            quarter = qq;
    }
    {
        uint32_t o = rec_qones(r.w[5], quarter);
        cnt += B ? o : quarter * 8 * kBs - o;
        p += rec_qbits(r.w[5], quarter);
    }
    uint64_t const x = rec_quarter(r.w[2], r.w[3], r.w[4], quarter); // the quarter's eight classes, six bits each
    uint32_t j = 0, k = 0, sp = 0;
    for (;; ++j)
    {
        k = (uint32_t)(x >> (6u * j)) & 63u;
        if (inv)
            k = kBs - k;
        sp = t->space[k];
        uint32_t c = B ? k : kBs - k;
        if (cnt + c >= i || j == 7)
            break;
        cnt += c;
        p += sp;
    }
    j += quarter << 3;
    uint32_t const target = (uint32_t)(i - cnt);
    uint32_t pos;
    if (k == 0 || k == kBs)
        pos = target - 1; // a uniform block: its target-th bit (of the only value it has)
    else
        pos = rrr_select_in_block<B>(t, k, sp ? read_int(v.btnr, p, sp) : 0, target, v.try_sparse != 0);
    return (begin * kK + j) * kBs + pos;
}

// ------------------------------------------------------------------------------------------------
// host side shared by rrr.cu and the CPU tests
// ------------------------------------------------------------------------------------------------
// C(n,k) for n,k <= 63 and the code lengths, computed once on the host (uploaded by rrr.cu, used by the CPU tests)
inline RrrTables const & host_tables()
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
        for (int n = 0; n < 63; ++n)
            for (int k = 0; k < 64; ++k)
            {
                if (n >= (int)kBinomSplit)
                    t.hi[n - kBinomSplit][k] = full[n][k];
                else
                    t.lo[n][k] = (uint32_t)full[n][k]; // C(33, 16) = 1.17e9 < 2^32
            }
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

// fills one 64-byte record from the REAL classes of its blocks (host and device share this)
__host__ __device__ inline void rrr_make_record(uint32_t const * k_real, uint32_t nblk_here, bool complete, uint8_t const * space, uint64_t ones_before,
                                                uint64_t bits_before, uint64_t * rec)
{
    bool inv = false;
    if (complete)
    { // only complete superblocks can be inverted (rrr_vector.hpp:203-228)
        uint32_t gt = 0;
        for (uint32_t j = 0; j < kK; ++j)
            gt += k_real[j] > kBs / 2;
        inv = gt > kK / 2;
    }
    uint64_t w[3] = {0, 0, 0};
    uint32_t ones = 0, bits = 0, qo[4] = {0, 0, 0, 0}, qb[4] = {0, 0, 0, 0};
    for (uint32_t j = 0; j < kK; ++j)
    {
        if ((j & 7) == 0)
        {
            qo[j >> 3] = ones;
            qb[j >> 3] = bits;
        }
        if (j >= nblk_here)
            continue;
        uint64_t c = inv ? kBs - k_real[j] : k_real[j];
        uint32_t bit = j * 6;
        w[bit >> 6] |= c << (bit & 63);
        if ((bit & 63) > 58)
            w[(bit >> 6) + 1] |= c >> (64 - (bit & 63));
        ones += k_real[j];
        bits += space[k_real[j]];
    }
    rec[0] = ones_before;
    rec[1] = bits_before | (inv ? kInvBit : 0);
    rec[2] = w[0];
    rec[3] = w[1];
    rec[4] = w[2];
    rec[5] = (uint64_t)qo[1] | ((uint64_t)qo[2] << 10) | ((uint64_t)qo[3] << 20) | ((uint64_t)qb[1] << 31) | ((uint64_t)qb[2] << 41) | ((uint64_t)qb[3] << 51);
    rec[6] = ones;
    rec[7] = bits;
}

// the 64-byte records of an rrr_vector<63> from the arrays the reference serialises (m_bt classes, m_rank, m_btnrp,
// m_invert: ingest, sdsl_format.cu); `rec` receives nsuper + 1 records, the last one holds the totals
inline void rrr_records_host(uint64_t const * bt_words /* packed 6-bit stored classes */, uint64_t nblocks, uint64_t nsuper, uint64_t ones,
                             std::vector<uint64_t> const & rank, std::vector<uint64_t> const & btnrp, std::vector<uint8_t> const & invert,
                             uint64_t total_bits_hint, std::vector<uint64_t> & rec)
{
    RrrTables const & t = host_tables();
    rec.assign(kRecWords * (nsuper + 1), 0);
    for (uint64_t g = 0; g < nsuper; ++g)
    {
        uint32_t k_real[kK];
        uint32_t here = 0;
        for (uint32_t j = 0; j < kK; ++j)
        {
            uint64_t b = g * kK + j;
            uint32_t c = 0;
            if (b < nblocks)
            {
                uint64_t pos = b * 6;
                uint64_t lo = bt_words[pos >> 6] >> (pos & 63);
                if ((pos & 63) > 58)
                    lo |= bt_words[(pos >> 6) + 1] << (64 - (pos & 63));
                c = (uint32_t)(lo & 63);
                if (invert[g])
                    c = kBs - c;
                ++here;
            }
            k_real[j] = c;
        }
        uint64_t * out = rec.data() + g * kRecWords;
        rrr_make_record(k_real, here, g * kK + kK <= nblocks, t.space, rank[g], btnrp[g], out);
        // keep the reference's invert bit verbatim (it equals the recomputed one for every complete superblock)
        out[1] = btnrp[g] | (invert[g] ? kInvBit : 0);
    }
    rec[kRecWords * nsuper] = ones;
    rec[kRecWords * nsuper + 1] = total_bits_hint;
}

} // namespace sdslgpu
