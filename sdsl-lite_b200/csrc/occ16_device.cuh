// occ16_device.cuh — the FM-index occurrence structure of this library (no SDSL counterpart; it answers the
// same csa.bwt.rank(i, c) / csa.bwt.inverse_select(i) the wavelet tree answers, suffix_array_helper.hpp:461-464,
// wt_pc.hpp:371-430, in 2 resp. 3 sector gathers instead of one per Huffman level).
//
// Why: on B200 a gather into a structure that is not L2-resident costs ~1/38 ns PER LOAD INSTRUCTION, whatever
// its width up to 32 bytes and whether or not its sector was just fetched (tools/probe_line.cu,
// profiles/r01_probe_line.txt).  A wt_huff<> rank over a byte alphabet is ~8 dependent gathers.  Spending HBM
// capacity instead: the comp-coded BWT symbol cc = 16*h + lo is indexed by 16 + 16 one-hot bitmaps in the
// 32-byte sector-block layout (count + 224 payload bits, common.cuh) —
//   level 0:  B0[h][i]  = [bwt[i] >> 4 == h]                                   (length n each)
//   level 1:  B1[lo][j] = [S[j] & 15 == lo], S = the symbols stably sorted by h  (length n each)
// so  rank(i, cc) = rank1(B1[lo], CH[h] + rank1(B0[h], i)) - rank1(B1[lo], CH[h]),  CH[h] = #symbols with a
// smaller high nibble: two gathers.  LF / inverse_select read bwt[i] (one byte) first: three gathers.
// sigma <= 16 needs level 0 only.  Cost: 32 * 8/7 bits + 1 byte per symbol (5.6 B) next to the 1.14 B/symbol
// wavelet tree; SDSLGPU_F_COMPACT keeps the tree alone.
#pragma once
#include "common.cuh"
#include "internal.h"

namespace sdslgpu
{

__device__ __forceinline__ uint64_t occ16_top(uint64_t const * __restrict__ top, uint64_t blk)
{
    // superblock 0 starts at count 0: texts below 3.7e9 symbols never touch the table
    return (blk >> kSuperShift) ? __ldg(top + (blk >> kSuperShift)) : 0ull;
}

__device__ __forceinline__ uint64_t occ16_rank1(bvblock const * __restrict__ blocks, uint64_t const * __restrict__ top, uint64_t pos)
{
    uint64_t blk = pos / kBlockBits;
    uint32_t rem = (uint32_t)(pos - blk * kBlockBits);
    uint32_t cnt, d[7];
    ld_block_half_line(blocks + blk, cnt, d);
    return occ16_top(top, blk) + cnt + block_prefix_popc(d, rem);
}

// rank1 at a <= b of the same bitmap; the usual case of a narrow interval is one gather for both
__device__ __forceinline__ void occ16_rank1_pair(bvblock const * __restrict__ blocks, uint64_t const * __restrict__ top, uint64_t & a, uint64_t & b)
{
    uint64_t ba = a / kBlockBits, bb = b / kBlockBits;
    uint32_t ra = (uint32_t)(a - ba * kBlockBits), rb = (uint32_t)(b - bb * kBlockBits);
    uint32_t cnt, d[7];
    ld_block_half_line(blocks + ba, cnt, d);
    uint64_t base = occ16_top(top, ba) + cnt;
    a = base + block_prefix_popc(d, ra);
    if (bb != ba)
    {
        ld_block_half_line(blocks + bb, cnt, d);
        base = occ16_top(top, bb) + cnt;
    }
    b = base + block_prefix_popc(d, rb);
}

// [a, b) -> the same half-open range of rows after prepending comp symbol cc:  l = a', r + 1 = b'
__device__ __forceinline__ void occ16_backward_step(Occ16Tab const * t, uint32_t cc, uint64_t & a, uint64_t & b)
{
    if (t->levels == 2)
    {
        uint32_t h = cc >> 4, lo = cc & 15u;
        occ16_rank1_pair(t->blocks[0][h], t->top[0][h], a, b);
        a += t->CH[h];
        b += t->CH[h];
        occ16_rank1_pair(t->blocks[1][lo], t->top[1][lo], a, b);
    }
    else
        occ16_rank1_pair(t->blocks[0][cc], t->top[0][cc], a, b);
    a += t->D[cc];
    b += t->D[cc];
}

// LF(i) = C[bwt[i]] + rank(i, bwt[i]) (suffix_array_helper.hpp:352-359); cc = comp code of bwt[i]
__device__ __forceinline__ uint64_t occ16_lf(Occ16Tab const * t, uint64_t i, uint32_t & cc)
{
    cc = __ldg(t->bwtc + i);
    uint64_t r;
    if (t->levels == 2)
    {
        uint32_t h = cc >> 4, lo = cc & 15u;
        r = t->CH[h] + occ16_rank1(t->blocks[0][h], t->top[0][h], i);
        r = occ16_rank1(t->blocks[1][lo], t->top[1][lo], r);
    }
    else
        r = occ16_rank1(t->blocks[0][cc], t->top[0][cc], i);
    return t->D[cc] + r;
}

} // namespace sdslgpu
