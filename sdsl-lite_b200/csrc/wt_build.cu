// wt_build.cu — the bit planes of wt_huff<> and wt_int<> built on the device (index CONSTRUCTION).
//
// Replaces the sequential fill of the reference's constructors (wt_pc.hpp:194-248 insert_char per symbol,
// wt_int.hpp:168-260 one stable partition per level) — 13.6 s on one core for the 2^28-byte tree of BASELINE
// config 4 — by one stable regrouping pass per tree depth:
//   wt_huff: nodes are numbered in BFS order and m_bv concatenates them in that order, so the bits of depth l are the
//            contiguous range [bv_pos(first node of depth l), bv_pos(first node of depth l+1)), and inside it the
//            symbols appear grouped by node, in text order.  The sequence grouped by "inner node at depth l" (symbols
//            whose code is already finished drop out) therefore yields the bits of depth l in order, and regrouped by
//            the nodes one depth further down it is the input of depth l+1.
//   wt_int:  level k of m_tree is the sequence stably sorted by its top k bits; bit = the next lower bit.
// The Huffman shape itself (<= 511 nodes) stays on the host (wt_shape.h build_huff_tree).
// wt_huff needs no sort at all: going one depth down, the elements of every node are split STABLY by their code bit —
// zeros to child 0, ones to child 1, elements whose code ends drop out — and where a child's elements start is known
// from the tree (bv_pos).  So an element's destination is "start of its child + number of elements of its node with the
// same bit before it", and that number is a rank query on the bits just written: one popcount pass + scan.cuh's prefix
// sum + one scatter pass per depth (wt_split_kernel), all hand-written.  wt_int (its nodes are not known in advance)
// sorts with the engine's own radix sort (radix.cuh; a library sort until the end of round 2: 0.084 - 0.17 s for 2^26
// values of 20 bits, now 0.080 s).
// The result is checked bit for bit against the host builders and the reference (tests/test_wt_gpu.py,
// tests/test_egress_gpu.py compare complete serialised trees).
#include "internal.h"
#include "radix.cuh"
#include "scan.cuh"
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

// per tree depth: the inner node every symbol sits in at that depth (relative to the depth's first node) and the bit
// its code has there
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

// ones per 32-bit word of the bit range just written: the input of the prefix sum behind wt_split_kernel
__global__ void __launch_bounds__(kThreads) wt_word_popc_kernel(uint32_t const * __restrict__ words32, uint64_t u0, uint64_t nwords, uint32_t * __restrict__ cnt)
{
    uint64_t const stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nwords; k += stride)
        cnt[k] = (uint32_t)__popc(words32[u0 + k]);
}

// the inner nodes of one tree depth, as the split needs them (indexed like LevelLut::key: relative to the depth's first node)
struct LevelNodes
{
    uint64_t seg_start[256]; // first element of the node in the depth's sequence
    uint64_t dst[2][256];    // where the elements going to child 0 / 1 start in the NEXT depth's sequence; kDropDst: a leaf
};
static constexpr uint64_t kDropDst = ~0ull;

// One depth down: element i of the current sequence (symbol c, inner node v = key[c], code bit b) moves to
// dst[b][v] + (number of elements of v with bit b before i).  The bits of the current depth already sit in the output
// vector at [start, start + cnt); `prefix` holds, per 32-bit word from word start / 32 on, the ones before that word.
__global__ void __launch_bounds__(kThreads) wt_split_kernel(uint8_t const * __restrict__ cur,
                                                            uint64_t cnt,
                                                            LevelLut const * __restrict__ lut,
                                                            LevelNodes const * __restrict__ nodes,
                                                            uint32_t nnodes,
                                                            uint32_t const * __restrict__ words32,
                                                            uint64_t const * __restrict__ prefix,
                                                            uint64_t start,
                                                            uint8_t * __restrict__ nxt)
{
    __shared__ uint16_t key[256];
    __shared__ uint32_t mask[8];
    __shared__ uint64_t seg_start[256], seg_ones[256], dst0[256], dst1[256];
    uint64_t const u0 = start >> 5;
    // ones in bits [u0 * 32, x) of the output vector
    auto ones_before = [&](uint64_t x) -> uint64_t {
        uint64_t const u = x >> 5;
        uint32_t const o = (uint32_t)(x & 31u);
        return prefix[u - u0] + (o ? (uint64_t)__popc(words32[u] & ((1u << o) - 1u)) : 0ull);
    };
    uint32_t const t = threadIdx.x; // kThreads == 256
    key[t] = lut->key[t];
    if (t < 8)
        mask[t] = lut->bit[t];
    uint64_t const base_ones = ones_before(start);
    if (t < nnodes)
    {
        seg_start[t] = nodes->seg_start[t];
        dst0[t] = nodes->dst[0][t];
        dst1[t] = nodes->dst[1][t];
        seg_ones[t] = ones_before(start + nodes->seg_start[t]) - base_ones;
    }
    __syncthreads();
    uint64_t const stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + t; i < cnt; i += stride)
    {
        uint32_t const c = cur[i], v = key[c], b = (mask[c >> 5] >> (c & 31u)) & 1u;
        uint64_t const r1 = ones_before(start + i) - base_ones - seg_ones[v]; // ones of node v before element i
        uint64_t const d = b ? dst1[v] : dst0[v];
        if (d != kDropDst)
            nxt[d + (b ? r1 : (i - seg_start[v]) - r1)] = (uint8_t)c;
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

    Buf bufa, bufb, luts, lnodes, wcnt, wpre, stmp;
    uint64_t const max_words32 = (n + 31) / 32 + 2; // a depth never has more than n bits
    if (bufa.alloc(n) != cudaSuccess || bufb.alloc(n) != cudaSuccess || luts.alloc((maxd + 1) * sizeof(LevelLut)) != cudaSuccess ||
        lnodes.alloc((maxd + 1) * sizeof(LevelNodes)) != cudaSuccess || wcnt.alloc(max_words32 * 4) != cudaSuccess ||
        wpre.alloc((max_words32 + 1) * 8) != cudaSuccess || stmp.alloc(scan_tmp_words(max_words32) * 8) != cudaSuccess)
    {
        cudaGetLastError();
        return SDSLGPU_ENOTSUP;
    }
    // per depth: the node (relative to the depth's first) and the code bit of every symbol; per inner node of the depth:
    // where its elements start and where those of its two children start one depth further down
    std::vector<LevelLut> lut(maxd + 1);
    std::vector<LevelNodes> nodes(maxd + 1);
    for (uint32_t l = 0; l <= maxd; ++l)
    {
        uint32_t const at_depth = first[l + 1] - first[l];
        if (at_depth > 256)
        {
            set_error("wt_huff device build: %u nodes at one depth", at_depth);
            return SDSLGPU_ECUDA;
        }
        for (int c = 0; c < 256; ++c)
            lut[l].key[c] = 0;
        std::memset(lut[l].bit, 0, sizeof(lut[l].bit));
        for (uint32_t j = 0; j < 256; ++j)
        {
            nodes[l].seg_start[j] = 0;
            nodes[l].dst[0][j] = nodes[l].dst[1][j] = kDropDst;
        }
        for (uint32_t v = first[l]; v < first[l + 1]; ++v)
            if (tree.child[v][0] != kWtUndef)
            {
                uint32_t const j = v - first[l];
                nodes[l].seg_start[j] = tree.bv_pos[v] - bv_start(first[l]);
                for (int b = 0; b < 2; ++b)
                {
                    uint32_t const w = tree.child[v][b];
                    if (tree.child[w][0] != kWtUndef) // the child is an inner node: its elements live on at depth l + 1
                        nodes[l].dst[b][j] = tree.bv_pos[w] - bv_start(first[l + 1]);
                }
            }
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
    SG_CUDA(cudaMemcpyAsync(lnodes.p, nodes.data(), nodes.size() * sizeof(LevelNodes), cudaMemcpyHostToDevice, s));

    uint8_t const * cur = d_text;
    uint8_t * spare[2] = {bufa.as<uint8_t>(), bufb.as<uint8_t>()};
    uint32_t * const words32 = reinterpret_cast<uint32_t *>(d_words);
    for (uint32_t l = 0; l < maxd; ++l)
    {
        uint64_t const start = bv_start(first[l]), m = bv_start(first[l + 1]) - start; // elements = bits of this depth
        if (m == 0)
            break;
        LevelLut const * dl = luts.as<LevelLut>() + l;
        wt_pack_huff_kernel<<<pack_grid(m), kThreads, 0, s>>>(cur, dl, m, start, words32);
        SG_CUDA(cudaGetLastError());
        uint64_t const next_m = bv_start(first[l + 2 <= maxd + 1 ? l + 2 : maxd + 1]) - bv_start(first[l + 1]);
        if (l + 1 >= maxd || next_m == 0)
            break; // the deepest inner nodes: nothing lives on below them
        // regroup for depth l + 1: stable split of every node by the bit just written
        uint64_t const u0 = start >> 5, nw = ((start + m + 31) >> 5) - u0;
        wt_word_popc_kernel<<<grid_for(nw), kThreads, 0, s>>>(words32, u0, nw, wcnt.as<uint32_t>());
        SG_CUDA(cudaGetLastError());
        SG_CUDA(exclusive_scan(wcnt.as<uint32_t>(), nw, wpre.as<uint64_t>(), stmp.as<uint64_t>(), s));
        uint8_t * nxt = spare[l & 1];
        wt_split_kernel<<<grid_for(m, 4), kThreads, 0, s>>>(cur, m, dl, lnodes.as<LevelNodes>() + l, first[l + 1] - first[l], words32, wpre.as<uint64_t>(), start,
                                                            nxt);
        SG_CUDA(cudaGetLastError());
        cur = nxt;
    }
    SG_CUDA(cudaStreamSynchronize(s));
    return SDSLGPU_OK;
}

// wt_int: d_seq (n values, device, overwritten) -> max_level, sigma and the level bits in d_words
// (*d_words_out is cudaMalloc'ed here: n * max_level bits + 2 words; the caller frees it)
int wt_int_planes_device(uint64_t * d_seq, uint64_t n, uint32_t * max_level_out, uint64_t * sigma_out, uint64_t ** d_words_out, cudaStream_t s)
{
    *d_words_out = nullptr;
    Buf scal, alt, sort_tmp;
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
    uint64_t const sort_bytes = radix_temp_bytes(n);
    uint64_t * d_words = nullptr;
    if (sort_tmp.alloc(sort_bytes) != cudaSuccess || cudaMalloc(reinterpret_cast<void **>(&d_words), nwords * 8) != cudaSuccess)
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
        uint32_t *no_vals = nullptr, *no_vals_alt = nullptr;
        SG_CUDA(radix_sort<false>(cur, nxt, no_vals, no_vals_alt, n, (int)shift, (int)levels, sort_tmp.p, s)); // `cur` = the result
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
