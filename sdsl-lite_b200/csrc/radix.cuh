// radix.cuh — stable least-significant-digit radix sort of u64 keys (optionally carrying u32 values), 8 bits per pass.
// CONSTRUCTION only (the suffix sorter of gpu_sa.cu, the level orders of wt_int in wt_build.cu); hand-written so that
// no library kernel runs anywhere in the engine (round 1 and most of round 2 used cub::DeviceRadixSort here; measured
// against it on the same builds, profiles/r02zz_bench_build.jsonl: csa_wt of a 2^28-byte text 0.60 - 0.98 -> 0.45 s,
// wt_int of 2^26 20-bit values 0.084 - 0.17 -> 0.080 s).
//
// One pass over a digit = three launches:
//   rs_hist_kernel      digit histogram of every tile of 4096 keys              hist[digit][tile]
//   exclusive_scan      (scan.cuh) over the digit-major table                   -> where each (digit, tile) group starts
//   rs_scatter_kernel   every CTA ranks the keys of its tile inside their digit, in memory order, and stores them
// Ranking: a warp owns 512 consecutive keys (16 per lane, striped, so loads coalesce); item by item the lanes that hold
// the same digit find each other with __match_any_sync, the lowest of them advances the warp's private counter of that
// digit by the size of the group, and a lane's rank is the old counter + the number of lower lanes in its group — no
// atomics, no barrier inside the loop.  One exclusive scan over the eight warps per digit turns the private counters
// into offsets.  Keys of one digit leave a tile as one contiguous run per digit (16 keys = 128 bytes on average).
#pragma once
#include "scan.cuh"

namespace sdslgpu
{

static constexpr int kRsThreads = 256;
static constexpr int kRsWarps = kRsThreads / 32;
static constexpr int kRsItems = 16; // keys per thread
static constexpr int kRsTile = kRsThreads * kRsItems;
static constexpr int kRsDigits = 256;

static_assert(kRsThreads == kRsDigits, "thread t owns digit t");

static __global__ void __launch_bounds__(kRsThreads) rs_hist_kernel(uint64_t const * __restrict__ keys, uint64_t n, uint32_t shift, uint32_t mask, uint64_t ntiles,
                                                              uint32_t * __restrict__ hist)
{
    __shared__ uint32_t cnt[kRsDigits];
    cnt[threadIdx.x] = 0;
    __syncthreads();
    uint64_t const base = (uint64_t)blockIdx.x * kRsTile;
#pragma unroll
    for (int i = 0; i < kRsItems; ++i)
    {
        uint64_t const p = base + (uint64_t)i * kRsThreads + threadIdx.x;
        if (p < n)
            atomicAdd(&cnt[(uint32_t)(keys[p] >> shift) & mask], 1u);
    }
    __syncthreads();
    hist[(uint64_t)threadIdx.x * ntiles + blockIdx.x] = cnt[threadIdx.x];
}

template <bool kPairs>
__global__ void __launch_bounds__(kRsThreads, 3) rs_scatter_kernel(uint64_t const * __restrict__ keys_in,
                                                                 uint32_t const * __restrict__ vals_in,
                                                                 uint64_t n,
                                                                 uint32_t shift,
                                                                 uint32_t mask,
                                                                 uint64_t ntiles,
                                                                 uint64_t const * __restrict__ offs,
                                                                 uint64_t * __restrict__ keys_out,
                                                                 uint32_t * __restrict__ vals_out)
{
    __shared__ uint32_t whist[kRsWarps][kRsDigits]; // per warp and digit: keys seen so far, later the warp's offset inside the tile's digit
    __shared__ uint64_t gbase[kRsDigits];           // where the tile's keys of a digit start in the output
    uint32_t const tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
    for (uint32_t k = tid; k < kRsWarps * kRsDigits; k += kRsThreads)
        (&whist[0][0])[k] = 0;
    gbase[tid] = offs[(uint64_t)tid * ntiles + blockIdx.x];
    __syncthreads();
    uint64_t const wbase = (uint64_t)blockIdx.x * kRsTile + (uint64_t)wid * (32 * kRsItems);
    uint64_t key[kRsItems];
    uint32_t rank2[kRsItems / 2]; // ranks inside the warp's 512 keys: two per register
#pragma unroll
    for (int i = 0; i < kRsItems; ++i)
    {
        uint64_t const p = wbase + (uint64_t)i * 32 + lane;
        key[i] = p < n ? keys_in[p] : 0ull;
    }
#pragma unroll
    for (int i = 0; i < kRsItems; ++i)
    {
        bool const valid = wbase + (uint64_t)i * 32 + lane < n;
        // lanes past the end get a value of their own: they match nobody
        uint32_t const d = valid ? ((uint32_t)(key[i] >> shift) & mask) : (uint32_t)kRsDigits + lane;
        uint32_t const peers = __match_any_sync(0xFFFFFFFFu, d);
        uint32_t const below = (uint32_t)__popc(peers & ((1u << lane) - 1u));
        uint32_t old = 0;
        if (valid && below == 0)
        { // the lowest lane of the group: the row is private to the warp, the groups of one item have different digits
            old = whist[wid][d];
            whist[wid][d] = old + (uint32_t)__popc(peers);
        }
        old = __shfl_sync(0xFFFFFFFFu, old, __ffs((int)peers) - 1);
        rank2[i >> 1] = (i & 1) ? (rank2[i >> 1] | ((old + below) << 16)) : (old + below);
        __syncwarp(); // the next item's group leaders read what this item's wrote
    }
    __syncthreads();
    { // digit tid: exclusive scan of the warps' counts
        uint32_t acc = 0;
#pragma unroll
        for (int w = 0; w < kRsWarps; ++w)
        {
            uint32_t const c = whist[w][tid];
            whist[w][tid] = acc;
            acc += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kRsItems; ++i)
    {
        uint64_t const p = wbase + (uint64_t)i * 32 + lane;
        if (p < n)
        {
            uint32_t const d = (uint32_t)(key[i] >> shift) & mask;
            uint64_t const pos = gbase[d] + whist[wid][d] + ((rank2[i >> 1] >> (16 * (i & 1))) & 0xFFFFu);
            keys_out[pos] = key[i];
            if (kPairs)
                vals_out[pos] = vals_in[p];
        }
    }
}

// device scratch of a sort of n keys
inline uint64_t radix_temp_bytes(uint64_t n)
{
    uint64_t const ntiles = (n + kRsTile - 1) / kRsTile, m = ntiles * kRsDigits;
    return m * 4 + 256 + (m + 1) * 8 + 256 + scan_tmp_words(m) * 8;
}

// Sorts by the bits [begin_bit, end_bit) of the keys, stably.  keys / keys_alt (and vals / vals_alt when kPairs) are two
// buffers of n entries; on return `keys` (and `vals`) point to the one that holds the result, the other is scratch.
template <bool kPairs>
inline cudaError_t radix_sort(uint64_t *& keys, uint64_t *& keys_alt, uint32_t *& vals, uint32_t *& vals_alt, uint64_t n, int begin_bit, int end_bit, void * temp,
                              cudaStream_t s)
{
    if (n == 0)
        return cudaSuccess;
    uint64_t const ntiles = (n + kRsTile - 1) / kRsTile, m = ntiles * kRsDigits;
    uint8_t * t = static_cast<uint8_t *>(temp);
    uint32_t * hist = reinterpret_cast<uint32_t *>(t);
    t += (m * 4 + 255) & ~255ull;
    uint64_t * offs = reinterpret_cast<uint64_t *>(t);
    t += ((m + 1) * 8 + 255) & ~255ull;
    uint64_t * scan_tmp = reinterpret_cast<uint64_t *>(t);
    for (int b = begin_bit; b < end_bit; b += 8)
    {
        int const bits = end_bit - b < 8 ? end_bit - b : 8;
        uint32_t const mask = (1u << bits) - 1u;
        rs_hist_kernel<<<(unsigned)ntiles, kRsThreads, 0, s>>>(keys, n, (uint32_t)b, mask, ntiles, hist);
        cudaError_t e = exclusive_scan(hist, m, offs, scan_tmp, s);
        if (e != cudaSuccess)
            return e;
        rs_scatter_kernel<kPairs><<<(unsigned)ntiles, kRsThreads, 0, s>>>(keys, vals, n, (uint32_t)b, mask, ntiles, offs, keys_alt, vals_alt);
        e = cudaGetLastError();
        if (e != cudaSuccess)
            return e;
        uint64_t * tk = keys;
        keys = keys_alt;
        keys_alt = tk;
        if (kPairs)
        {
            uint32_t * tv = vals;
            vals = vals_alt;
            vals_alt = tv;
        }
    }
    return cudaSuccess;
}

} // namespace sdslgpu
