// sd_device.cuh — device-side view of an sd_vector<> image and its per-query primitives (rank_1, select_1), used by
// sd.cu's kernels and batch ops; like bv_device.cuh / rrr_device.cuh it compiles as plain C++ under SDSLGPU_HOST_EMU
// for tests/test_device_logic_cpu.py.
#pragma once
#include "bv_device.cuh"

namespace sdslgpu
{

struct SdView
{
    uint64_t size, m;
    uint32_t wl;
    BvView high;
    uint64_t const * low; // m entries of wl bits, packed like int_vector<0>
    // select_0 samples: samp0[q] = sector block of `high` in which the (q * 2^log_s0 + 1)-th ZERO of the vector is
    // "crossed" (see sd_select0_one); nullptr: not built
    uint32_t const * samp0;
    uint32_t log_s0;
};

__device__ __forceinline__ uint64_t sd_low(SdView const & v, uint64_t j)
{
    return read_int(v.low, j * v.wl, v.wl);
}

// bit `o` (0..223) of a block's payload held in registers, without dynamic register indexing
__device__ __forceinline__ uint32_t sd_reg_bit(uint32_t const (&d)[7], uint32_t o)
{
    uint32_t w = 0;
#pragma unroll
    for (uint32_t j = 0; j < 7; ++j)
        w = (j == (o >> 5)) ? d[j] : w;
    return (w >> (o & 31u)) & 1u;
}

// number of ones in [0, i)   (sd_vector.hpp:553-575): the zero that closes bucket i >> wl in `high`, then backwards over
// the bucket's elements while their low part is >= i's.  The select hands out the sector block it ended in, so the
// bits below the zero are tested in registers; only a bucket that continues into the previous block costs a gather
// per element (the reference reads `high` bit by bit, :566-573).
__device__ __forceinline__ uint64_t sd_rank1_one(SdView const & v, uint64_t i)
{
    uint64_t hv = i >> v.wl;
    uint32_t d[7];
    uint64_t g;
    uint64_t sh = bv_select<0>(v.high, hv + 1, d, g); // end of bucket hv in `high`
    uint64_t rl = sh - hv;                          // elements with high part <= hv
    if (rl == 0)
        return 0;
    uint64_t const vl = i & ((1ull << v.wl) - 1), start = g * kBlockBits;
    do
    {
        if (!sh)
            return 0;
        --sh;
        --rl;
    } while ((sh >= start ? sd_reg_bit(d, (uint32_t)(sh - start)) : bv_bit(v.high, sh)) && sd_low(v, rl) >= vl);
    return rl + 1;
}

// position of the i-th one, 1 <= i <= m   (sd_vector.hpp:621-630)
__device__ __forceinline__ uint64_t sd_select1_one(SdView const & v, uint64_t i)
{
    return sd_low(v, i - 1) + ((bv_select<1>(v.high, i) + 1 - i) << v.wl);
}

// ------------------------------------------------------------------------------------------------
// select_0 in O(1) expected gathers (the job of select_0_support_sd, sd_vector.hpp:752-921; same answers as the binary
// search of select_support_sd<0>, :637-663).
//
// `high` lists, bucket by bucket (bucket h = positions [h * 2^wl, (h+1) * 2^wl) of the vector), one 1 per element of the
// bucket and then a 0.  For a prefix high[0, j) let W(j) = zeros(j) * 2^wl - ones(j).  Right after the zero that closes
// bucket h, W = (h+1) * 2^wl - (ones of the vector before the end of bucket h) = number of ZEROS of the vector in
// buckets 0..h =: V(h).  W only rises at zeros, so the bucket holding the i-th zero of the vector ends at the first
// position j with W(j) >= i ("the crossing").  One sector block of `high` carries the count of ones before it, hence
// W at its start; inside the block W follows from word popcounts.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int64_t sd_w_at(SdView const & v, uint64_t pos_in_high, uint64_t ones_before)
{
    return (int64_t)((pos_in_high - ones_before) << v.wl) - (int64_t)ones_before;
}

// W right after the last zero of block g (v_last; only if the block has a zero) and right after the last zero before
// the block (v_prev; 0 if there is none): the block "owns" the zeros i of the vector with v_prev < i <= v_last.
// Build-time helper of the sample table (sd.cu) — and of the host image in tests/cpp/sd_on_host.cpp.
__device__ __forceinline__ bool sd_block_zero_span(SdView const & v, uint64_t g, int64_t & v_prev, int64_t & v_last)
{
    uint32_t cnt, d[7];
    ld_block(v.high.blocks + g, cnt, d);
    uint64_t const r1 = __ldg(v.high.top + (g >> kSuperShift)) + cnt, start = g * kBlockBits;
    int64_t w = sd_w_at(v, start, r1);
    // the run of ones that ends right before the block belongs to a bucket that is still open: undo it
    uint64_t run = 0;
    while (start > run && bv_bit(v.high, start - 1 - run))
        ++run;
    v_prev = (start > run) ? w + (int64_t)run : 0;
    int64_t const step = (int64_t)(1ull << v.wl) + 1;
    bool any = false;
    v_last = v_prev;
    uint64_t const valid = v.high.nbits > start ? v.high.nbits - start : 0; // bits past the end of `high` are padding
#pragma unroll
    for (uint32_t k = 0; k < 7; ++k)
    {
        uint32_t x = ~d[k];
        if (valid < 32ull * (k + 1))
            x &= valid > 32ull * k ? (uint32_t)((1ull << (valid - 32ull * k)) - 1ull) : 0u;
        uint32_t const z = (uint32_t)__popc(x);
        int64_t const w_end = w + (int64_t)z * step - 32;
        if (z)
        {
            any = true;
            v_last = w_end + (int64_t)__clz((int)x); // + the ones above the word's last zero = W right behind that zero
        }
        w = w_end;
    }
    return any;
}

// select_support_sd<0>::select (sd_vector.hpp:637-663): binary search over select_1 for the last one with fewer than i
// zeros before it.  O(log m) selects; kept as the fallback of sd_select0_one and for handles without samples.
__device__ __forceinline__ uint64_t sd_select0_bsearch(SdView const & v, uint64_t i)
{
    uint64_t lb = 1, rb = v.m + 1, r0 = 0, pos = ~0ull;
    while (lb < rb)
    {
        uint64_t mid = lb + (rb - lb) / 2;
        uint64_t x = sd_select1_one(v, mid);
        uint64_t rank0 = x + 1 - mid;
        if (rank0 >= i)
            rb = mid;
        else
        {
            r0 = rank0;
            pos = x;
            lb = mid + 1;
        }
    }
    return pos + i - r0;
}

static constexpr uint32_t kSdSelect0Walk = 8; // sector blocks of `high` looked at before giving up on the samples

// position of the i-th zero of the vector, 1 <= i <= size - m, from the samples: one sample gather, the block(s) of
// `high` up to the crossing (1.3 on random data), and one low part per element of the crossed bucket that lies behind
// the answer (0.6 on average).  Returns ~0 when the crossing is not within kSdSelect0Walk blocks of the sample
// (long runs of ones in clustered vectors): the caller then uses sd_select0_bsearch.
__device__ __forceinline__ uint64_t sd_select0_one(SdView const & v, uint64_t i)
{
    int64_t const need = (int64_t)i, step = (int64_t)(1ull << v.wl) + 1;
    uint64_t g = __ldg(v.samp0 + ((i - 1) >> v.log_s0));
    uint64_t const last_block = v.high.nbits / kBlockBits;
    for (uint32_t walked = 0; walked < kSdSelect0Walk && g <= last_block; ++walked, ++g)
    {
        uint32_t cnt, d[7];
        ld_block(v.high.blocks + g, cnt, d);
        uint64_t const r1 = __ldg(v.high.top + (g >> kSuperShift)) + cnt, start = g * kBlockBits;
        int64_t w = sd_w_at(v, start, r1), w_word = 0;
        uint32_t zb = 0, zb_word = 0, dw = 0;
        int32_t kw = -1;
#pragma unroll
        for (int32_t k = 0; k < 7; ++k)
        { // the first word whose last zero has W >= i behind it holds the crossing
            uint32_t const x = ~d[k], z = (uint32_t)__popc(x);
            int64_t const w_end = w + (int64_t)z * step - 32;
            if (kw < 0 && z && w_end + (int64_t)__clz((int)x) >= need)
            {
                kw = k;
                w_word = w;
                zb_word = zb;
                dw = d[k];
            }
            w = w_end;
            zb += z;
        }
        if (kw < 0)
            continue;
        uint32_t x = ~dw, t = 0, pos = 0;
        for (;;)
        { // zeros of the word in order; the last one satisfies the test, so this terminates
            pos = (uint32_t)__ffs((int)x) - 1u;
            x &= x - 1u;
            ++t;
            if (w_word + (int64_t)t * step - (int64_t)(pos + 1u) >= need)
                break;
        }
        uint32_t const bit = 32u * (uint32_t)kw + pos;
        uint64_t const e = start + bit;                      // the zero of `high` that closes the bucket of the answer
        uint64_t k1 = r1 + bit - (zb_word + t - 1u);         // ones of `high` before e = elements in buckets <= that bucket
        uint64_t const base = (e - k1) << v.wl;              // first position of the bucket (e - k1 zeros precede e)
        // the elements of the bucket are the run of ones right below e; the answer is i - 1 + (elements before it)
        uint32_t run = 0;
        if (pos)
        {
            uint32_t const below = dw << (32u - pos); // the pos bits below the zero, moved to the top
            run = (uint32_t)__clz((int)~below);
            run = run < pos ? run : pos;
        }
        for (uint32_t u = 0; u < run; ++u, --k1)
            if (base + sd_low(v, k1 - 1) < i - 1 + k1)
                return i - 1 + k1;
        if (run == pos)
        { // the run may go on below the word (a bucket with many elements): bit by bit
            uint64_t ph = e - run;
            while (ph > 0 && bv_bit(v.high, ph - 1))
            {
                if (base + sd_low(v, k1 - 1) < i - 1 + k1)
                    break;
                --k1;
                --ph;
            }
        }
        return i - 1 + k1;
    }
    return ~0ull;
}

} // namespace sdslgpu
