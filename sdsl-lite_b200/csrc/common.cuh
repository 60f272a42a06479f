// common.cuh — device-side primitives shared by every kernel.
//
// Device equivalents of the reference's word-level helpers (bits.hpp:486-502 cnt, :586-612 sel,
// :194-233 lo_set/lo_unset, :777-790 read_int) plus the sector-block layout the B200 rank/select
// structures are built on.
#pragma once
#include <cstdint>
#ifdef SDSLGPU_HOST_EMU
// tests only (tests/cpp/device_on_host.cpp): the per-query device functions of this header and of bv_device.cuh
// compiled as plain C++, so that their logic is checked against the oracle on a box without a GPU.  The product
// library never defines SDSLGPU_HOST_EMU.
#include <cstring>
#define __device__
#define __host__
#define __forceinline__ inline
#define __align__(n) alignas(n)
struct uint2
{
    uint32_t x, y;
};
static inline int __popc(uint32_t x)
{
    return __builtin_popcount(x);
}
static inline int __popcll(uint64_t x)
{
    return __builtin_popcountll(x);
}
static inline int __clz(int x)
{
    return x ? __builtin_clz((unsigned)x) : 32;
}
static inline int __ffs(int x)
{
    return __builtin_ffs(x);
}
static inline uint64_t __umul64hi(uint64_t a, uint64_t b)
{
    return (uint64_t)(((unsigned __int128)a * b) >> 64);
}
template <class T>
static inline T __ldg(T const * p)
{
    return *p;
}
#else
#include <cuda_runtime.h>
#endif

namespace sdslgpu
{

// ------------------------------------------------------------------------------------------------
// The B200-native bit-vector block: ONE 32-byte DRAM sector per block.
//   cnt  : number of 1-bits in [first bit of the enclosing superblock, first bit of this block)
//   d[7] : 224 payload bits, bit i of the vector = (d[(i%224)>>5] >> (i&31)) & 1   (LSB first, the
//          same bit order as int_vector<1>, int_vector.hpp:1900-1904)
// A superblock is 2^24 blocks (3.76e9 bits), so cnt always fits 32 bits; the absolute count of the
// superblock lives in a tiny u64 table (`top`) that stays in L1/L2.
// One rank query = one 32-byte sector gather (the reference needs two cache lines: a 16-byte table
// pair and an 8-byte data word, rank_support_v.hpp:133-135).
// ------------------------------------------------------------------------------------------------
struct __align__(32) bvblock
{
    uint32_t cnt;
    uint32_t d[7];
};
static constexpr uint32_t kBlockBits = 224;
static constexpr uint32_t kSuperShift = 24; // blocks per superblock = 1 << 24

#ifdef SDSLGPU_HOST_EMU
inline void ld_block(bvblock const * p, uint32_t & cnt, uint32_t (&d)[7])
{
    cnt = p->cnt;
    std::memcpy(d, p->d, sizeof(d));
}
inline void ld_block_half_line(bvblock const * p, uint32_t & cnt, uint32_t (&d)[7])
{
    ld_block(p, cnt, d);
}
inline uint32_t ld_nc_u32(uint32_t const * p)
{
    return *p;
}
inline uint64_t ld_nc_u64(uint64_t const * p)
{
    return *p;
}
#else
// 256-bit (one sector) read-only gather: a single LDG.E.256 on sm_100a.
__device__ __forceinline__ void ld_block(bvblock const * p, uint32_t & cnt, uint32_t (&d)[7])
{
    asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(cnt), "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6])
                 : "l"(p));
}

// same gather, but tells the L2 to fill only the 64-byte half line on a miss (SASS: LDG...LTC64B.256).  A B200 L2
// miss otherwise fills the whole 128-byte line (ncu: 4 DRAM sectors per 32-byte gather, profiles/r01_probe_fetch*);
// for single-block lookups (rank, one wavelet-tree level) the other 96 bytes are never used.
__device__ __forceinline__ void ld_block_half_line(bvblock const * p, uint32_t & cnt, uint32_t (&d)[7])
{
    asm volatile("ld.global.nc.L1::no_allocate.L2::64B.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(cnt), "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6])
                 : "l"(p));
}

// 128-bit read-only gather (SDSL-layout rank table pair: absolute count + 7x9-bit relative counts)
__device__ __forceinline__ void ld_pair(uint64_t const * p, uint64_t & a, uint64_t & b)
{
    asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p));
}

__device__ __forceinline__ uint64_t ld_nc_u64(uint64_t const * p)
{
    uint64_t v;
    asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ uint32_t ld_nc_u32(uint32_t const * p)
{
    uint32_t v;
    asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
// streaming (evict-first) accesses for the query / result streams: they are touched exactly once
__device__ __forceinline__ uint64_t ld_stream_u64(uint64_t const * p)
{
    uint64_t v;
    asm volatile("ld.global.cs.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_stream_u64(uint64_t * p, uint64_t v)
{
    asm volatile("st.global.cs.u64 [%0], %1;" ::"l"(p), "l"(v));
}

__device__ __forceinline__ uint32_t ld_stream_u32(uint32_t const * p)
{
    uint32_t v;
    asm volatile("ld.global.cs.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ uint32_t ld_stream_u16(uint16_t const * p)
{
    uint16_t v;
    asm volatile("ld.global.cs.u16 %0, [%1];" : "=h"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_stream_u32(uint32_t * p, uint32_t v)
{
    asm volatile("st.global.cs.u32 [%0], %1;" ::"l"(p), "r"(v));
}

__device__ __forceinline__ void st_stream(uint32_t * p, uint32_t v)
{
    st_stream_u32(p, v);
}
__device__ __forceinline__ void st_stream(uint64_t * p, uint64_t v)
{
    st_stream_u64(p, v);
}
#endif // SDSLGPU_HOST_EMU

// lo_set[k] for 0 <= k <= 63 (bits.hpp:194-211); k == 64 is never needed on the device paths
__device__ __forceinline__ uint64_t lo_set64(uint32_t k)
{
    return (1ULL << k) - 1ULL;
}

// position (0-based) of the k-th (1-based) set bit of a 32-bit word, 1 <= k <= popc(x):
// popcount bisection, 5 steps (device form of bits::sel, bits.hpp:586-612)
__device__ __forceinline__ uint32_t sel32(uint32_t x, uint32_t k)
{
    uint32_t pos = 0, c;
    c = __popc(x & 0xFFFFu);
    if (k > c)
    {
        k -= c;
        pos += 16;
        x >>= 16;
    }
    c = __popc(x & 0xFFu);
    if (k > c)
    {
        k -= c;
        pos += 8;
        x >>= 8;
    }
    c = __popc(x & 0xFu);
    if (k > c)
    {
        k -= c;
        pos += 4;
        x >>= 4;
    }
    c = __popc(x & 0x3u);
    if (k > c)
    {
        k -= c;
        pos += 2;
        x >>= 2;
    }
    pos += (k > (x & 1u)) ? 1u : 0u;
    return pos;
}

__device__ __forceinline__ uint32_t sel64(uint64_t x, uint32_t k)
{
    uint32_t lo = (uint32_t)x, c = __popc(lo);
    return (k <= c) ? sel32(lo, k) : 32u + sel32((uint32_t)(x >> 32), k - c);
}

// number of 1-bits among the first `rem` (0..223) payload bits of a block
__device__ __forceinline__ uint32_t block_prefix_popc(uint32_t const (&d)[7], uint32_t rem)
{
    uint32_t w = rem >> 5, o = rem & 31u, r = 0;
    uint32_t part = (1u << o) - 1u;
#pragma unroll
    for (uint32_t j = 0; j < 7; ++j)
    {
        uint32_t m = (j < w) ? 0xFFFFFFFFu : ((j == w) ? part : 0u);
        r += __popc(d[j] & m);
    }
    return r;
}

template <int B>
__device__ __forceinline__ uint32_t block_popc(uint32_t const (&d)[7])
{
    uint32_t r = 0;
#pragma unroll
    for (uint32_t j = 0; j < 7; ++j)
        r += __popc(B ? d[j] : ~d[j]);
    return r;
}

// position within the block (0..223) of the k-th (1-based) B-bit; requires 1 <= k <= block_popc<B>
template <int B>
__device__ __forceinline__ uint32_t block_select(uint32_t const (&d)[7], uint32_t k)
{
    uint32_t pos = 0;
    uint32_t x = B ? d[0] : ~d[0];
#pragma unroll
    for (uint32_t j = 0; j < 6; ++j)
    {
        uint32_t c = __popc(x);
        if (k > c)
        {
            k -= c;
            pos += 32;
            x = B ? d[j + 1] : ~d[j + 1];
        }
        else
            break;
    }
    return pos + sel32(x, k);
}

// unaligned read of `len` (1..64) bits at absolute bit position `pos` from packed 64-bit words
// (device form of bits::read_int, bits.hpp:777-790).  The arrays are padded with one extra word.
__device__ __forceinline__ uint64_t read_int(uint64_t const * __restrict__ d, uint64_t pos, uint32_t len)
{
    uint64_t const * w = d + (pos >> 6);
    uint32_t off = (uint32_t)(pos & 63);
    uint64_t lo = __ldg(w) >> off;
    if (off + len > 64)
        lo |= __ldg(w + 1) << (64 - off);
    return len == 64 ? lo : (lo & ((1ULL << len) - 1ULL));
}

} // namespace sdslgpu
