// fan.cuh — device side of the result fan-out of the multi-GPU group calls (internal.h: Fan; host side: group.cu).
// A kernel that has just produced answer v for local query index p also hands it to the other members of the group:
//   fan.width == 0   u64 peer stores: dst[r][p] = v, r < fan.n (dst = the peers' result arrays, SDSLGPU_GATHER_FUSED)
//   fan.width == w   the answers cross NVLink as w-bit fields (SDSLGPU_GATHER_PACKED): dst[r] is the region of the
//                    peer's staging buffer that belongs to THIS member, field p at bits [p*w, (p+1)*w) — the layout of
//                    int_vector<w> (bits::write_int, bits.hpp:737-760).  A warp holds 32 consecutive answers = 32*w bits
//                    = w/2 whole words (w even), so every word is assembled in registers by shuffles and stored once,
//                    without atomics; the receiver widens the fields again (group.cu fan_unpack_kernel).
#pragma once
#include "internal.h"

namespace sdslgpu
{

static constexpr uint32_t kPackMinWidth = 22; // <= 4 fields overlap one 64-bit word
static constexpr uint32_t kPackMaxWidth = 62;

__device__ __forceinline__ void fan_store(Fan const & fan, uint64_t p, uint64_t v)
{
    for (uint32_t r = 0; r < fan.n; ++r)
        fan.dst[r][p] = v;
}

// Called by all 32 lanes of a warp: lane holds the answer of local index p0 + lane (p0 a multiple of 32); lanes
// >= nvalid hold none.  SDSLGPU_NPOS becomes the all-ones field.
__device__ __forceinline__ void fan_store_packed(Fan const & fan, uint64_t p0, uint32_t lane, uint32_t nvalid, uint64_t v)
{
    uint32_t const w = fan.width;
    v = lane < nvalid ? (v & ((1ull << w) - 1ull)) : 0ull;
    uint32_t const bit0 = 64u * lane; // this lane assembles bits [bit0, bit0 + 64) of the warp's 32 * w bits
    uint32_t const f0 = bit0 / w;     // the first field that reaches into them
    uint64_t acc = 0;
#pragma unroll
    for (uint32_t t = 0; t < 4; ++t)
    {
        uint32_t const f = f0 + t;
        uint64_t const val = __shfl_sync(0xFFFFFFFFu, v, (int)(f & 31u));
        int32_t const sh = (int32_t)(f * w) - (int32_t)bit0; // start of field f relative to the word
        if (f < 32u && sh < 64)
            acc |= sh >= 0 ? val << sh : val >> (-sh);
    }
    if (lane < w / 2 && bit0 < nvalid * w)
    {
        uint64_t const word = (p0 >> 5) * (w / 2) + lane;
        for (uint32_t r = 0; r < fan.n; ++r)
            fan.dst[r][word] = acc;
    }
}

// words of a staging region that holds s fields of width w written by fan_store_packed
__host__ __device__ __forceinline__ uint64_t fan_region_words(uint64_t s, uint32_t w)
{
    return ((s + 31) >> 5) * (w / 2);
}

} // namespace sdslgpu
