// wt_device.cuh — per-query wavelet-tree primitives over a tree staged in shared memory, shared by wt.cu
// (wt_huff queries) and fm.cu (backward search, LF walks).
#pragma once
#include "bits_access.cuh"
#ifdef SDSLGPU_HOST_EMU
#include "wt_tree.h" // tests/cpp/wt_on_host.cpp: the per-query functions below as plain C++
#else
#include "internal.h"
#endif

namespace sdslgpu
{

#ifndef SDSLGPU_HOST_EMU
// cooperative 16-byte copy of the node table + paths (~16 KB) into shared memory, once per CTA
__device__ __forceinline__ void stage_tree(WtTree const * __restrict__ g, WtTree * s)
{
    uint4 const * src = reinterpret_cast<uint4 const *>(g);
    uint4 * dst = reinterpret_cast<uint4 *>(s);
    for (uint32_t k = threadIdx.x; k < sizeof(WtTree) / 16; k += blockDim.x)
        dst[k] = __ldg(src + k);
    __syncthreads();
}
#endif

// rank(i, c) for one query (wt_pc.hpp:371-399): path_len dependent sector gathers on the concatenated m_bv
template <class Bits>
__device__ __forceinline__ uint64_t wt_rank_one(Bits const & bits, WtTree const * t, uint64_t sigma, uint64_t i, uint32_t c)
{
    if (t->c_to_leaf[c] == kWtUndef)
        return 0;
    if (sigma == 1)
        return i;
    uint64_t p = t->path[c];
    uint32_t len = (uint32_t)(p >> 56);
    uint64_t r = i;
    uint32_t v = 0;
    for (uint32_t l = 0; l < len && r; ++l, p >>= 1)
    {
        uint64_t o = bits.rank1(t->bv_pos[v] + r) - t->bv_pos_rank[v];
        r = (p & 1) ? o : r - o;
        v = t->child[v][p & 1];
    }
    return r;
}

// (rank(i, wt[i]), wt[i]) (wt_pc.hpp:411-430): per level ONE sector yields both the bit and the rank
template <class Bits>
__device__ __forceinline__ uint64_t wt_inverse_select_one(Bits const & bits, WtTree const * t, uint64_t i, uint32_t & sym)
{
    uint32_t v = 0;
    while (t->child[v][0] != kWtUndef)
    {
        uint32_t bit;
        uint64_t o = bits.rank1_and_bit(t->bv_pos[v] + i, bit) - t->bv_pos_rank[v];
        i = bit ? o : i - o;
        v = t->child[v][bit];
    }
    sym = (uint32_t)t->bv_pos_rank[v];
    return i;
}

} // namespace sdslgpu
