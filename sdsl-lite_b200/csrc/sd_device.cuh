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
};

__device__ __forceinline__ uint64_t sd_low(SdView const & v, uint64_t j)
{
    return read_int(v.low, j * v.wl, v.wl);
}

// number of ones in [0, i)   (sd_vector.hpp:553-575)
__device__ __forceinline__ uint64_t sd_rank1_one(SdView const & v, uint64_t i)
{
    uint64_t hv = i >> v.wl;
    uint64_t sh = bv_select<0>(v.high, hv + 1); // end of bucket hv in `high`
    uint64_t rl = sh - hv;                      // elements with high part <= hv
    if (rl == 0)
        return 0;
    uint64_t vl = i & ((1ull << v.wl) - 1);
    do
    {
        if (!sh)
            return 0;
        --sh;
        --rl;
    } while (bv_bit(v.high, sh) && sd_low(v, rl) >= vl);
    return rl + 1;
}

// position of the i-th one, 1 <= i <= m   (sd_vector.hpp:621-630)
__device__ __forceinline__ uint64_t sd_select1_one(SdView const & v, uint64_t i)
{
    return sd_low(v, i - 1) + ((bv_select<1>(v.high, i) + 1 - i) << v.wl);
}

} // namespace sdslgpu
