// gpu_sa.cu — suffix array, BWT and SA samples on the device (index CONSTRUCTION, not the query hot path).
//
// Replaces the construction step the reference performs with its vendored divsufsort on one CPU thread
// (construct_sa.hpp:103-137, construct_bwt.hpp:53-60, csa_sampling_strategy.hpp:98-115): minutes for a 1 GiB text
// (SURVEY.md §7.3-5).  Here: prefix doubling.  Round 0 sorts the suffixes by their first 8 bytes (one 64-bit radix
// sort); every later round sorts by (rank of the h-prefix, rank of the h-prefix h positions later) and doubles h,
// until all ranks are distinct.  Random text finishes after round 0; natural-language text in a handful of rounds;
// the worst case (a^n) needs log2(n) rounds.  The sort itself is the engine's own stable LSD radix sort (radix.cuh; round 1
// and most of round 2 called a library sort here: 0.60 - 0.98 s for a 2^28-byte text end to end, now 0.45 s).  Any
// correct suffix sorter yields the same array, and the resulting index is checked byte for byte against the reference's.
#include "internal.h"
#include "radix.cuh"
#include "scan.cuh"

namespace sdslgpu
{

__global__ void __launch_bounds__(kThreads) sa_init_keys_kernel(uint8_t const * __restrict__ t, uint64_t n, uint64_t * __restrict__ key, uint32_t * __restrict__ val)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    uint64_t k = 0;
#pragma unroll
    for (int b = 0; b < 8; ++b)
        k = (k << 8) | t[i + b]; // the text buffer is zero padded past the sentinel
    key[i] = k;
    val[i] = (uint32_t)i;
}

__global__ void __launch_bounds__(kThreads) sa_flag_heads_kernel(uint64_t const * __restrict__ key, uint64_t n, uint32_t * __restrict__ flag)
{
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n)
        return;
    flag[j] = (j == 0 || key[j] != key[j - 1]) ? 1u : 0u;
}

// rank of the suffix at sorted position j = (number of group heads in [0, j]) - 1; scattered to text order
__global__ void __launch_bounds__(kThreads)
    sa_scatter_rank_kernel(uint64_t const * __restrict__ heads_before, uint32_t const * __restrict__ flag, uint32_t const * __restrict__ val, uint64_t n, uint32_t * __restrict__ rank)
{
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n)
        return;
    rank[val[j]] = (uint32_t)(heads_before[j] + flag[j] - 1);
}

__global__ void __launch_bounds__(kThreads) sa_next_keys_kernel(uint32_t const * __restrict__ rank, uint32_t const * __restrict__ val, uint64_t n, uint64_t h, uint64_t * __restrict__ key)
{
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n)
        return;
    uint64_t i = val[j];
    uint64_t second = (i + h < n) ? (uint64_t)rank[i + h] + 1 : 0;
    key[j] = ((uint64_t)rank[i] << 32) | second;
}

__global__ void __launch_bounds__(kThreads)
    sa_bwt_kernel(uint8_t const * __restrict__ t,
                  uint32_t const * __restrict__ sa,
                  uint64_t n,
                  uint32_t dens,
                  uint32_t isa_dens,
                  uint8_t * __restrict__ bwt,
                  uint64_t * __restrict__ samples,
                  uint64_t * __restrict__ isa_samples)
{
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n)
        return;
    uint32_t p = sa[j];
    bwt[j] = p ? t[p - 1] : t[n - 1]; // construct_bwt.hpp:53-60
    if (j % dens == 0)
        samples[j / dens] = p; // csa_sampling_strategy.hpp:98-115
    if (p % isa_dens == 0)
        isa_samples[p / isa_dens] = j; // csa_sampling_strategy.hpp:758-779
}

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
    T * as()
    {
        return static_cast<T *>(p);
    }
};
} // namespace

// text_host: len zero-free bytes.  Outputs (host): bwt[len+1], samples[ceil((len+1)/dens)].
// Returns SDSLGPU_ENOTSUP when the text is too long for 32-bit suffix indices or the device lacks memory, so the
// caller can fall back to the host SA-IS builder.
int gpu_suffix_array_bwt(uint8_t const * text_host,
                         uint64_t len,
                         uint32_t dens,
                         uint32_t isa_dens,
                         std::vector<uint8_t> & bwt,
                         std::vector<uint64_t> & samples,
                         std::vector<uint64_t> & isa_samples,
                         uint32_t * rounds_out,
                         cudaStream_t s)
{
    uint64_t n = len + 1;
    if (n >= (1ull << 32) - 1)
        return SDSLGPU_ENOTSUP;
    Buf t, k0, k1, v0, v1, rk, fl, hb, tmp, sort_tmp, dbwt, dsamp;
    uint64_t nsamp = (n + dens - 1) / dens, nisa = (n - 1) / isa_dens + 1;
    uint64_t const sort_bytes = radix_temp_bytes(n);
    if (t.alloc(n + 16) != cudaSuccess || k0.alloc(n * 8) != cudaSuccess || k1.alloc(n * 8) != cudaSuccess || v0.alloc(n * 4) != cudaSuccess ||
        v1.alloc(n * 4) != cudaSuccess || rk.alloc(n * 4) != cudaSuccess || fl.alloc(n * 4) != cudaSuccess || hb.alloc((n + 1) * 8) != cudaSuccess ||
        tmp.alloc(scan_tmp_words(n) * 8) != cudaSuccess || sort_tmp.alloc(sort_bytes) != cudaSuccess || dsamp.alloc(nisa * 8) != cudaSuccess)
    {
        cudaGetLastError();
        return SDSLGPU_ENOTSUP;
    }
    SG_CUDA(cudaMemsetAsync(t.p, 0, n + 16, s));
    if (len)
        SG_CUDA(cudaMemcpyAsync(t.p, text_host, len, cudaMemcpyHostToDevice, s));
    // two key and two value buffers; k_cur / v_cur hold the current order, the others are the sort's scratch
    uint64_t *k_cur = k0.as<uint64_t>(), *k_alt = k1.as<uint64_t>();
    uint32_t *v_cur = v0.as<uint32_t>(), *v_alt = v1.as<uint32_t>();
    sa_init_keys_kernel<<<blocks_for(n), kThreads, 0, s>>>(t.as<uint8_t>(), n, k_cur, v_cur);
    SG_CUDA(cudaGetLastError());
    uint32_t rounds = 0;
    int rank_bits = 1;
    while ((1ull << rank_bits) <= n + 1)
        ++rank_bits;
    for (uint64_t h = 8;; h <<= 1)
    {
        // round 0 sorts by the 8-byte prefix; later rounds only need the bits the two ranks occupy
        int hi_bit = rounds == 0 ? 64 : 32 + rank_bits;
        SG_CUDA(radix_sort<true>(k_cur, k_alt, v_cur, v_alt, n, 0, hi_bit, sort_tmp.p, s));
        ++rounds;
        sa_flag_heads_kernel<<<blocks_for(n), kThreads, 0, s>>>(k_cur, n, fl.as<uint32_t>());
        SG_CUDA(cudaGetLastError());
        SG_CUDA(exclusive_scan(fl.as<uint32_t>(), n, hb.as<uint64_t>(), tmp.as<uint64_t>(), s));
        uint64_t groups = 0;
        SG_CUDA(cudaMemcpyAsync(&groups, hb.as<uint64_t>() + n, 8, cudaMemcpyDeviceToHost, s));
        SG_CUDA(cudaStreamSynchronize(s));
        if (groups == n || h >= n)
            break;
        sa_scatter_rank_kernel<<<blocks_for(n), kThreads, 0, s>>>(hb.as<uint64_t>(), fl.as<uint32_t>(), v_cur, n, rk.as<uint32_t>());
        SG_CUDA(cudaGetLastError());
        sa_next_keys_kernel<<<blocks_for(n), kThreads, 0, s>>>(rk.as<uint32_t>(), v_cur, n, h, k_cur);
        SG_CUDA(cudaGetLastError());
    }
    if (rounds_out)
        *rounds_out = rounds;
    // BWT + samples; the key buffers are free again: reuse one for the BWT bytes and the scan output for the samples
    uint8_t * d_bwt = reinterpret_cast<uint8_t *>(k_alt);
    uint64_t * d_samp = hb.as<uint64_t>();
    sa_bwt_kernel<<<blocks_for(n), kThreads, 0, s>>>(t.as<uint8_t>(), v_cur, n, dens, isa_dens, d_bwt, d_samp, dsamp.as<uint64_t>());
    SG_CUDA(cudaGetLastError());
    bwt.resize(n);
    samples.resize(nsamp);
    isa_samples.resize(nisa);
    SG_CUDA(cudaMemcpyAsync(isa_samples.data(), dsamp.p, nisa * 8, cudaMemcpyDeviceToHost, s));
    SG_CUDA(cudaMemcpyAsync(bwt.data(), d_bwt, n, cudaMemcpyDeviceToHost, s));
    SG_CUDA(cudaMemcpyAsync(samples.data(), d_samp, nsamp * 8, cudaMemcpyDeviceToHost, s));
    SG_CUDA(cudaStreamSynchronize(s));
    return SDSLGPU_OK;
}

} // namespace sdslgpu
