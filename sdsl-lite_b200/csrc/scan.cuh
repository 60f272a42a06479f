// scan.cuh — device-wide exclusive prefix sum (u32 counts -> u64 offsets), used by every builder
// (rank directories, select samples, CSR outputs).  Three phases: per-tile reduce, single-CTA scan of
// the tile totals, per-tile rescan with the tile base.  Hand-written (no CUB) so the build path has
// no library dependency; it runs once per structure and is HBM-streaming bound.
#pragma once
#include "common.cuh"

namespace sdslgpu
{

static constexpr int kScanThreads = 256;
static constexpr int kScanItems = 8; // per thread
static constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ uint64_t warp_incl_scan(uint64_t v)
{
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        uint64_t t = __shfl_up_sync(0xFFFFFFFFu, v, o);
        if ((threadIdx.x & 31) >= o)
            v += t;
    }
    return v;
}

// exclusive scan of one value per thread across the CTA (blockDim.x <= 1024); returns the exclusive
// prefix and writes the CTA total to `total`
__device__ __forceinline__ uint64_t block_excl_scan(uint64_t v, uint64_t & total)
{
    __shared__ uint64_t warp_tot[32];
    __shared__ uint64_t all;
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    uint64_t inc = warp_incl_scan(v);
    if (lane == 31)
        warp_tot[wid] = inc;
    __syncthreads();
    if (wid == 0)
    {
        uint64_t t = (lane < nw) ? warp_tot[lane] : 0;
        uint64_t ti = warp_incl_scan(t);
        if (lane < nw)
            warp_tot[lane] = ti - t;
        if (lane == 31)
            all = ti;
    }
    __syncthreads();
    uint64_t r = warp_tot[wid] + inc - v;
    total = all;
    __syncthreads();
    return r;
}

template <class In>
__global__ void __launch_bounds__(kScanThreads) scan_tile_sums_kernel(In const * __restrict__ in, uint64_t n, uint64_t * __restrict__ tile_sum)
{
    uint64_t base = (uint64_t)blockIdx.x * kScanTile;
    uint64_t s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k)
    {
        uint64_t i = base + (uint64_t)k * kScanThreads + threadIdx.x;
        if (i < n)
            s += in[i];
    }
    uint64_t tot;
    block_excl_scan(s, tot);
    if (threadIdx.x == 0)
        tile_sum[blockIdx.x] = tot;
}

// in-place exclusive scan of `m` u64 values by ONE CTA; total -> *grand_total
static __global__ void __launch_bounds__(1024) scan_single_cta_kernel(uint64_t * __restrict__ v, uint64_t m, uint64_t * __restrict__ grand_total)
{
    uint64_t carry = 0;
    for (uint64_t base = 0; base < m; base += blockDim.x)
    {
        uint64_t i = base + threadIdx.x;
        uint64_t x = (i < m) ? v[i] : 0;
        uint64_t tot;
        uint64_t e = block_excl_scan(x, tot);
        if (i < m)
            v[i] = carry + e;
        carry += tot;
    }
    if (threadIdx.x == 0 && grand_total)
        *grand_total = carry;
}

template <class In>
__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(In const * __restrict__ in, uint64_t n, uint64_t const * __restrict__ tile_base, uint64_t * __restrict__ out)
{
    // thread t owns kScanItems CONSECUTIVE items so that its local prefix is a simple running sum
    uint64_t base = (uint64_t)blockIdx.x * kScanTile + (uint64_t)threadIdx.x * kScanItems;
    uint64_t x[kScanItems];
    uint64_t s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k)
    {
        uint64_t i = base + k;
        x[k] = (i < n) ? (uint64_t)in[i] : 0;
        s += x[k];
    }
    uint64_t tot;
    uint64_t e = block_excl_scan(s, tot) + tile_base[blockIdx.x];
#pragma unroll
    for (int k = 0; k < kScanItems; ++k)
    {
        uint64_t i = base + k;
        if (i < n)
            out[i] = e;
        e += x[k];
    }
}

// out[i] = sum_{j<i} in[j] for i in [0, n]  (out has n+1 entries: out[n] = grand total)
// `tile_tmp` must hold ceil(n / kScanTile) + 1 u64 values.
template <class In>
inline cudaError_t exclusive_scan(In const * in, uint64_t n, uint64_t * out, uint64_t * tile_tmp, cudaStream_t s)
{
    uint64_t tiles = (n + kScanTile - 1) / kScanTile;
    if (n == 0)
    {
        return cudaMemsetAsync(out, 0, 8, s);
    }
    scan_tile_sums_kernel<In><<<(unsigned)tiles, kScanThreads, 0, s>>>(in, n, tile_tmp);
    scan_single_cta_kernel<<<1, 1024, 0, s>>>(tile_tmp, tiles, out + n);
    scan_apply_kernel<In><<<(unsigned)tiles, kScanThreads, 0, s>>>(in, n, tile_tmp, out);
    return cudaGetLastError();
}

inline uint64_t scan_tmp_words(uint64_t n)
{
    return (n + kScanTile - 1) / kScanTile + 1;
}

} // namespace sdslgpu
