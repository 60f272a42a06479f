// bits_access.cuh — the bit-vector "policy" the wavelet-tree and FM-index kernels are instantiated with, the
// device analogue of the reference's t_bitvector / t_rank / t_select template parameters (wt_pc.hpp:53-58):
//   PlainBits : bit_vector + rank_support_v + select_support_mcl   -> sector blocks (bv_device.cuh)
//   RrrBits   : rrr_vector<63> + rank_support_rrr + select_support_rrr -> fused records (rrr_device.cuh),
//               i.e. wt_huff<rrr_vector<63>> / csa_wt<wt_huff<rrr_vector<63>>>
#pragma once
#include <cstddef>

#include "bv_device.cuh"
#include "rrr_device.cuh"

namespace sdslgpu
{

struct PlainBits
{
    BvView v;
    static constexpr size_t kSmem = 0;
    __device__ __forceinline__ void attach(unsigned char *)
    {}
    __device__ __forceinline__ uint64_t rank1(uint64_t pos) const
    {
        return bv_rank1(v, pos);
    }
    __device__ __forceinline__ uint64_t rank1_and_bit(uint64_t pos, uint32_t & bit) const
    {
        return bv_rank1_and_bit(v, pos, bit);
    }
    template <int B>
    __device__ __forceinline__ uint64_t select(uint64_t i) const
    {
        return bv_select<B>(v, i);
    }
};

struct RrrBits
{
    RrrView v;
    RrrTables const * t; // shared-memory copy of the binomials, set by attach()
    static constexpr size_t kSmem = sizeof(RrrTables);
    __device__ __forceinline__ void attach(unsigned char * smem)
    {
#ifdef SDSLGPU_HOST_EMU
        (void)smem;
        t = v.tables;
#else
        RrrTables * s = reinterpret_cast<RrrTables *>(smem);
        stage_rrr(v.tables, s);
        t = s;
#endif
    }
    __device__ __forceinline__ uint64_t rank1(uint64_t pos) const
    {
        return rrr_rank1_one(v, t, pos);
    }
    __device__ __forceinline__ uint64_t rank1_and_bit(uint64_t pos, uint32_t & bit) const
    {
        return rrr_rank1_and_bit(v, t, pos, bit);
    }
    template <int B>
    __device__ __forceinline__ uint64_t select(uint64_t i) const
    {
        return rrr_select_one<B>(v, t, i);
    }
};

} // namespace sdslgpu
