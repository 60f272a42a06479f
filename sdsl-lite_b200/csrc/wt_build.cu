// wt_build.cu — the bit planes of wt_huff<> and wt_int<> built on the device (index CONSTRUCTION).
//
// Replaces the sequential fill of the reference's constructors (wt_pc.hpp:194-248 insert_char per symbol,
// wt_int.hpp:168-260 one stable partition per level) — 13.6 s on one core for the 2^28-byte tree of BASELINE
// config 4 — by one stable radix pass per tree depth:
//   wt_huff: nodes are numbered in BFS order and m_bv concatenates them in that order, so the bits of depth l are the
//            contiguous range [bv_pos(first node of depth l), bv_pos(first node of depth l+1)), and inside it the
//            symbols appear grouped by node, in text order.  Sorting the sequence stably by "inner node at depth l"
//            (symbols whose code is already finished sort behind everything and drop out) therefore yields the bits
//            of depth l in order; the sorted sequence is the input of depth l+1.
//   wt_int:  level k of m_tree is the sequence stably sorted by its top k bits; bit = the next lower bit.
// The Huffman shape itself (<= 511 nodes) stays on the host (wt.cu build_huff_tree).  Sorting primitive:
// cub::DeviceRadixSort (CCCL, shipped with the toolkit) as in gpu_sa.cu — builder only, no query kernel uses it.
// The result is checked bit for bit against the host builders and the reference (tests/test_wt_gpu.py,
// tests/test_egress_gpu.py compare complete serialised trees).
#include <cub/device/device_radix_sort.cuh>

#include "internal.h"
#include "wt_device.cuh"

namespace sdslgpu
{

namespace
{

struct Buf
{
    void * p = nullptr;
    ~Buf()
    {
        if (p)
            cudaFree(p);
    }
    cudaError_t alloc(uint64_t bytes)
    {
        return cudaMalloc(&p, bytes ? bytes : 8);
    }
    template <class T>
    T * as() const
    {
        return static_cast<T *>(p);
    }
};

// per tree depth: sort key of every symbol (inner node at that depth, relative to the depth's first node; kDrop
// position = "code finished") and the bit its code has there
struct LevelLut
{
    uint16_t key[256];
    uint32_t bit[8];
};

__global__ void __launch_bounds__(kThreads) wt_hist_kernel(uint8_t const * __restrict__ text, uint64_t n, unsigned long long * __restrict__ hist)
{
    __shared__ unsigned int h[256];
    h[threadIdx.x] = 0; // kThreads == 256
    __syncthreads();
    uint64_t const stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        atomicAdd(&h[text[i]], 1u);
    __syncthreads();
    if (h[threadIdx.x])
        atomicAdd(&hist[threadIdx.x], (unsigned long long)h[threadIdx.x]);
}

__global__ void __launch_bounds__(kThreads) wt_level_keys_kernel(uint8_t const * __restrict__ vals, uint64_t cnt, LevelLut const * __restrict__ lut, uint16_t * __restrict__ keys)
{
    __shared__ uint16_t key[256];
    key[threadIdx.x] = lut->key[threadIdx.x];
    __syncthreads();
    uint64_t const stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < cnt; k += stride)
        keys[k] = key[vals[k]];
}

// bits [start, start + m) of the output vector = bit of the m first elements of the sorted sequence.  One warp
// per 32-bit output word (ballot); words shared with the neighbouring depth are merged with atomicOr (the output is
// zero-initialised).
__global__ void __launch_bounds__(kThreads)
    wt_pack_huff_kernel(uint8_t const * __restrict__ vals, LevelLut const * __restrict__ lut, uint64_t m, uint64_t start, uint32_t * __restrict__ out32)
{
    uint32_t mask[8];
#pragma unroll
    for (int j = 0; j < 8; ++j)
        mask[j] = lut->bit[j];
    uint32_t const lane = threadIdx.x & 31u;
    uint64_t const u0 = start >> 5, u1 = (start + m + 31) >> 5;
    uint64_t const nwarps = (uint64_t)gridDim.x * (kThreads / 32);
    for (uint64_t u = u0 + (uint64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); u < u1; u += nwarps)
    {
        uint64_t b = (u << 5) + lane;
        uint32_t bit = 0;
        if (b >= start && b < start + m)
        {
            uint32_t c = vals[b - start];
            bit = (mask[c >> 5] >> (c & 31u)) & 1u;
        }
        uint32_t word = __ballot_sync(0xFFFFFFFFu, bit);
        if (lane == 0 && word)
            atomicOr(out32 + u, word);
    }
}

__global__ void __launch_bounds__(kThreads)
    wt_pack_int_kernel(uint64_t const * __restrict__ cur, uint32_t shift, uint64_t m, uint64_t start, uint32_t * __restrict__ out32)
{
    uint32_t const lane = threadIdx.x & 31u;
    uint64_t const u0 = start >> 5, u1 = (start + m + 31) >> 5;
    uint64_t const nwarps = (uint64_t)gridDim.x * (kThreads / 32);
    for (uint64_t u = u0 + (uint64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); u < u1; u += nwarps)
    {
        uint64_t b = (u << 5) + lane;
        uint32_t bit = (b >= start && b < start + m) ? (uint32_t)((cur[b - start] >> shift) & 1u) : 0u;
        uint32_t word = __ballot_sync(0xFFFFFFFFu, bit);
        if (lane == 0 && word)
            atomicOr(out32 + u, word);
    }
}

__global__ void __launch_bounds__(kThreads) wt_max_kernel(uint64_t const * __restrict__ seq, uint64_t n, unsigned long long * __restrict__ out)
{
    uint64_t const stride = (uint64_t)gridDim.x * blockDim.x;
    unsigned long long mx = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        mx = seq[i] > mx ? seq[i] : mx;
#pragma unroll
    for (int d = 16; d; d >>= 1)
    {
        unsigned long long o = __shfl_xor_sync(0xFFFFFFFFu, mx, d);
        mx = o > mx ? o : mx;
    }
    if ((threadIdx.x & 31u) == 0)
        atomicMax(out, mx);
}

// number of distinct values of a sorted sequence
__global__ void __launch_bounds__(kThreads) wt_distinct_kernel(uint64_t const * __restrict__ sorted, uint64_t n, unsigned long long * __restrict__ out)
{
    uint64_t const stride = (uint64_t)gridDim.x * blockDim.x;
    unsigned int c = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        c += (i == 0 || sorted[i] != sorted[i - 1]) ? 1u : 0u;
#pragma unroll
    for (int d = 16; d; d >>= 1)
        c += __shfl_xor_sync(0xFFFFFFFFu, c, d);
    if ((threadIdx.x & 31u) == 0 && c)
        atomicAdd(out, (unsigned long long)c);
}

unsigned pack_grid(uint64_t m)
{
    return grid_for(((m + 31) >> 5) * 32 + 32);
}

} // namespace

// symbol counts of a device-resident text
int wt_histogram_device(uint8_t const * d_text, uint64_t n, uint64_t (&C)[256], cudaStream_t s)
{
    Buf hist;
    if (hist.alloc(256 * 8) != cudaSuccess)
    {
        cudaGetLastError();
        return SDSLGPU_ENOTSUP;
    }
    SG_CUDA(cudaMemsetAsync(hist.p, 0, 256 * 8, s));
    if (n)
        wt_hist_kernel<<<grid_for(n, 16), kThreads, 0, s>>>(d_text, n, hist.as<unsigned long long>());
    SG_CUDA(cudaGetLastError());
    SG_CUDA(cudaMemcpyAsync(C, hist.p, 256 * 8, cudaMemcpyDeviceToHost, s));
    SG_CUDA(cudaStreamSynchronize(s));
    return SDSLGPU_OK;
}

// d_words: (bits + 63) / 64 + 2 words; receives m_bv of the tree `tree` (shape already fixed on the host) over the
// device-resident text.  SDSLGPU_ENOTSUP = no device memory for the scratch: the caller falls back to the host fill.
int wt_huff_planes_device(uint8_t const * d_text, uint64_t n, WtTree const & tree, uint64_t bits, uint64_t * d_words, cudaStream_t s)
{
    SG_CUDA(cudaMemsetAsync(d_words, 0, (((bits + 63) >> 6) + 2) * 8, s));
    uint32_t const nn = tree.nnodes;
    if (n == 0 || bits == 0 || nn < 3)
        return SDSLGPU_OK;
    // depth ranges of the BFS numbering
    std::vector<uint32_t> depth(nn, 0);
    uint32_t maxd = 0;
    for (uint32_t v = 0; v < nn; ++v)
        if (tree.child[v][0] != kWtUndef)
        {
            depth[tree.child[v][0]] = depth[tree.child[v][1]] = depth[v] + 1;
            maxd = depth[v] + 1 > maxd ? depth[v] + 1 : maxd;
        }
    std::vector<uint32_t> first(maxd + 2, nn);
    for (uint32_t v = nn; v-- > 0;)
        first[depth[v]] = v;
    first[maxd + 1] = nn;
    auto bv_start = [&](uint32_t v) { return v < nn ? tree.bv_pos[v] : bits; };

    Buf bufa, bufb, k0, k1, luts, cubtmp;
    size_t cub_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (uint16_t const *)nullptr, (uint16_t *)nullptr, (uint8_t const *)nullptr, (uint8_t *)nullptr, n, 0, 16, s);
    if (bufa.alloc(n) != cudaSuccess || bufb.alloc(n) != cudaSuccess || k0.alloc(n * 2) != cudaSuccess || k1.alloc(n * 2) != cudaSuccess ||
        luts.alloc((maxd + 1) * sizeof(LevelLut)) != cudaSuccess || cubtmp.alloc(cub_bytes) != cudaSuccess)
    {
        cudaGetLastError();
        return SDSLGPU_ENOTSUP;
    }
    // sort keys / code bits of every symbol at every depth
    std::vector<LevelLut> lut(maxd + 1);
    std::vector<uint32_t> drop(maxd + 1);
    for (uint32_t l = 0; l <= maxd; ++l)
    {
        drop[l] = first[l + 1] - first[l];
        for (int c = 0; c < 256; ++c)
            lut[l].key[c] = (uint16_t)drop[l];
        std::memset(lut[l].bit, 0, sizeof(lut[l].bit));
    }
    for (int c = 0; c < 256; ++c)
    {
        if (tree.c_to_leaf[c] == kWtUndef)
            continue;
        uint64_t p = tree.path[c];
        uint32_t len = (uint32_t)(p >> 56), v = 0;
        for (uint32_t l = 0; l < len; ++l, p >>= 1)
        {
            lut[l].key[c] = (uint16_t)(v - first[l]);
            if (p & 1)
                lut[l].bit[c >> 5] |= 1u << (c & 31);
            v = tree.child[v][p & 1];
        }
    }
    SG_CUDA(cudaMemcpyAsync(luts.p, lut.data(), lut.size() * sizeof(LevelLut), cudaMemcpyHostToDevice, s));

    uint8_t const * cur = d_text;
    uint8_t * spare[2] = {bufa.as<uint8_t>(), bufb.as<uint8_t>()};
    uint64_t cnt = n;
    for (uint32_t l = 0; l < maxd; ++l)
    {
        uint64_t start = bv_start(first[l]), m = bv_start(first[l + 1]) - start;
        if (m == 0)
            break;
        LevelLut const * dl = luts.as<LevelLut>() + l;
        if (l > 0)
        { // regroup by inner node of this depth; finished codes drop behind the first m elements
            int key_bits = 1;
            while ((1u << key_bits) <= drop[l])
                ++key_bits;
            wt_level_keys_kernel<<<grid_for(cnt, 4), kThreads, 0, s>>>(cur, cnt, dl, k0.as<uint16_t>());
            SG_CUDA(cudaGetLastError());
            uint8_t * nxt = spare[l & 1];
            size_t need = 0;
            cub::DeviceRadixSort::SortPairs(nullptr, need, k0.as<uint16_t>(), k1.as<uint16_t>(), cur, nxt, cnt, 0, key_bits, s);
            if (need > cub_bytes)
            {
                set_error("wt_huff device build: radix-sort scratch grew from %llu to %llu bytes", (unsigned long long)cub_bytes, (unsigned long long)need);
                return SDSLGPU_ECUDA;
            }
            SG_CUDA(cub::DeviceRadixSort::SortPairs(cubtmp.p, need, k0.as<uint16_t>(), k1.as<uint16_t>(), cur, nxt, cnt, 0, key_bits, s));
            cur = nxt;
        }
        wt_pack_huff_kernel<<<pack_grid(m), kThreads, 0, s>>>(cur, dl, m, start, reinterpret_cast<uint32_t *>(d_words));
        SG_CUDA(cudaGetLastError());
        cnt = m;
    }
    SG_CUDA(cudaStreamSynchronize(s));
    return SDSLGPU_OK;
}

// wt_int: d_seq (n values, device, overwritten) -> max_level, sigma and the level bits in d_words
// (*d_words_out is cudaMalloc'ed here: n * max_level bits + 2 words; the caller frees it)
int wt_int_planes_device(uint64_t * d_seq, uint64_t n, uint32_t * max_level_out, uint64_t * sigma_out, uint64_t ** d_words_out, cudaStream_t s)
{
    *d_words_out = nullptr;
    Buf scal, alt, cubtmp;
    if (scal.alloc(16) != cudaSuccess || alt.alloc(n * 8) != cudaSuccess)
    {
        cudaGetLastError();
        return SDSLGPU_ENOTSUP;
    }
    SG_CUDA(cudaMemsetAsync(scal.p, 0, 16, s));
    wt_max_kernel<<<grid_for(n, 8), kThreads, 0, s>>>(d_seq, n, scal.as<unsigned long long>());
    SG_CUDA(cudaGetLastError());
    uint64_t max_elem = 0;
    SG_CUDA(cudaMemcpyAsync(&max_elem, scal.p, 8, cudaMemcpyDeviceToHost, s));
    SG_CUDA(cudaStreamSynchronize(s));
    uint32_t hi = 0;
    for (uint64_t x = max_elem ? max_elem : 1; x >>= 1;)
        ++hi;
    uint32_t const levels = hi + 1; // wt_int.hpp:182 (an all-zero sequence still gets one level)
    uint64_t const bits = n * levels, nwords = ((bits + 63) >> 6) + 2;
    size_t cub_bytes = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, cub_bytes, (uint64_t const *)nullptr, (uint64_t *)nullptr, n, 0, 64, s);
    uint64_t * d_words = nullptr;
    if (cubtmp.alloc(cub_bytes) != cudaSuccess || cudaMalloc(reinterpret_cast<void **>(&d_words), nwords * 8) != cudaSuccess)
    {
        cudaGetLastError();
        return SDSLGPU_ENOTSUP;
    }
    Buf words_guard;
    words_guard.p = d_words; // freed on every error path; released to the caller at the end
    SG_CUDA(cudaMemsetAsync(d_words, 0, nwords * 8, s));
    uint64_t * cur = d_seq;
    uint64_t * nxt = alt.as<uint64_t>();
    for (uint32_t k = 0; k < levels; ++k)
    {
        uint32_t shift = levels - k - 1;
        wt_pack_int_kernel<<<pack_grid(n), kThreads, 0, s>>>(cur, shift, n, (uint64_t)k * n, reinterpret_cast<uint32_t *>(d_words));
        SG_CUDA(cudaGetLastError());
        // next level's order: stable by the top k+1 bits
        size_t need = 0;
        cub::DeviceRadixSort::SortKeys(nullptr, need, cur, nxt, n, (int)shift, (int)levels, s);
        if (need > cub_bytes)
        {
            set_error("wt_int device build: radix-sort scratch grew from %llu to %llu bytes", (unsigned long long)cub_bytes, (unsigned long long)need);
            return SDSLGPU_ECUDA;
        }
        SG_CUDA(cub::DeviceRadixSort::SortKeys(cubtmp.p, need, cur, nxt, n, (int)shift, (int)levels, s));
        uint64_t * t = cur;
        cur = nxt;
        nxt = t;
    }
    wt_distinct_kernel<<<grid_for(n, 8), kThreads, 0, s>>>(cur, n, scal.as<unsigned long long>() + 1);
    SG_CUDA(cudaGetLastError());
    uint64_t sigma = 0;
    SG_CUDA(cudaMemcpyAsync(&sigma, scal.as<unsigned long long>() + 1, 8, cudaMemcpyDeviceToHost, s));
    SG_CUDA(cudaStreamSynchronize(s));
    *max_level_out = levels;
    *sigma_out = sigma;
    *d_words_out = d_words;
    words_guard.p = nullptr;
    return SDSLGPU_OK;
}

} // namespace sdslgpu
