// sd.cu — sd_vector<> (Elias-Fano): device-side builder and batched rank / select / access.
//
// Replaces (results bit-exact; m_low and m_high are bit-identical to the reference's):
//   sd_vector ctor              sd_vector.hpp:218-257 -> sd_count / sd_scatter kernels + scan, then the
//                                                         sector-block rank/select image over `high`
//   rank_support_sd::rank       sd_vector.hpp:553-575 -> sd_rank_kernel
//   select_support_sd::select   sd_vector.hpp:621-664 -> sd_select_kernel<B>
//   sd_vector::operator[]       sd_vector.hpp:328-349 -> sd_access_kernel
// `high` (m ones, 2^logm zeros: always about half dense) gets the same one-sector rank blocks and select
// samples as a plain bit vector; select_support_mcl<1>/<0> of the reference (sd_vector.hpp:162-163) are
// therefore served by bv_select<1>/<0>.
#include "binned.cuh"
#include "internal.h"
#include "scan.cuh"
#include "sd_device.cuh"

namespace sdslgpu
{

__global__ void __launch_bounds__(kThreads) sd_rank_kernel(SdView const v, int b, uint64_t const * __restrict__ idx, uint64_t n, uint64_t * __restrict__ out)
{
    uint64_t const stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += stride)
    {
        uint64_t i = ld_stream_u64(idx + q);
        uint64_t r = SDSLGPU_NPOS;
        if (i <= v.size)
        {
            r = sd_rank1_one(v, i);
            if (!b)
                r = i - r;
        }
        st_stream_u64(out + q, r);
    }
}

__global__ void __launch_bounds__(kThreads) sd_access_kernel(SdView const v, uint64_t const * __restrict__ idx, uint64_t n, uint64_t * __restrict__ out)
{
    uint64_t const stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += stride)
    {
        uint64_t i = ld_stream_u64(idx + q);
        uint64_t r = SDSLGPU_NPOS;
        if (i < v.size)
        { // sd_vector.hpp:328-349: walk bucket hv backwards until low <= vl
            uint64_t hv = i >> v.wl;
            uint64_t sh = bv_select<0>(v.high, hv + 1);
            uint64_t rl = sh - hv;
            r = 0;
            if (rl)
            {
                uint64_t vl = i & ((1ull << v.wl) - 1);
                --sh;
                --rl;
                bool alive = true;
                while (bv_bit(v.high, sh) && sd_low(v, rl) > vl)
                {
                    if (sh == 0)
                    {
                        alive = false;
                        break;
                    }
                    --sh;
                    --rl;
                }
                r = (alive && bv_bit(v.high, sh) && sd_low(v, rl) == vl) ? 1 : 0;
            }
        }
        st_stream_u64(out + q, r);
    }
}

template <int B>
__global__ void __launch_bounds__(kThreads) sd_select_kernel(SdView const v, uint64_t const * __restrict__ idx, uint64_t n, uint64_t * __restrict__ out)
{
    uint64_t const args = B ? v.m : v.size - v.m;
    uint64_t const stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += stride)
    {
        uint64_t i = ld_stream_u64(idx + q);
        uint64_t r = SDSLGPU_NPOS;
        if (i >= 1 && i <= args)
        {
            if (B)
                r = sd_select1_one(v, i);
            else
            { // binary search over select_1 for the last one with fewer than i zeros before it (sd_vector.hpp:637-663)
                uint64_t lb = 1, rb = v.m + 1, r0 = 0, pos = ~0ull;
                while (lb < rb)
                {
                    uint64_t mid = lb + (rb - lb) / 2;
                    uint64_t x = sd_select1_one(v, mid);
                    uint64_t rank0 = x + 1 - mid;
                    if (rank0 >= i)
                        rb = mid;
                    else
                    {
                        r0 = rank0;
                        pos = x;
                        lb = mid + 1;
                    }
                }
                r = pos + i - r0;
            }
        }
        st_stream_u64(out + q, r);
    }
}

// ------------------------------------------------------------------------------------------------
// builder
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) sd_count_kernel(uint64_t const * __restrict__ words, uint64_t nbits, uint64_t nwords, uint32_t * __restrict__ cnt)
{
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nwords)
        return;
    uint64_t x = words[k];
    if (k == nwords - 1 && (nbits & 63))
        x &= (1ull << (nbits & 63)) - 1;
    cnt[k] = __popcll(x);
}

// one thread per input word: its j-th one (globally the (before + j)-th) writes low[idx] = pos mod 2^wl and
// sets bit (pos >> wl) + idx of `high`   (sd_vector.hpp:233-252, without the sequential cursor)
__global__ void __launch_bounds__(kThreads) sd_scatter_kernel(uint64_t const * __restrict__ words,
                                                              uint64_t nbits,
                                                              uint64_t nwords,
                                                              uint64_t const * __restrict__ before,
                                                              uint32_t wl,
                                                              unsigned long long * __restrict__ low,
                                                              unsigned long long * __restrict__ high)
{
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nwords)
        return;
    uint64_t x = words[k];
    if (k == nwords - 1 && (nbits & 63))
        x &= (1ull << (nbits & 63)) - 1;
    uint64_t j = before[k];
    uint64_t const mask = (1ull << wl) - 1;
    while (x)
    {
        uint32_t o = __ffsll((long long)x) - 1;
        x &= x - 1;
        uint64_t pos = k * 64 + o;
        uint64_t lp = j * wl, lv = pos & mask;
        uint32_t lo = (uint32_t)(lp & 63);
        if (lv)
        {
            atomicOr(low + (lp >> 6), (unsigned long long)(lv << lo));
            if (lo + wl > 64)
                atomicOr(low + (lp >> 6) + 1, (unsigned long long)(lv >> (64 - lo)));
        }
        uint64_t hp = (pos >> wl) + j;
        atomicOr(high + (hp >> 6), 1ull << (hp & 63));
        ++j;
    }
}

static SdView sd_view(SdImage const & d)
{
    SdView v;
    v.size = d.size;
    v.m = d.m;
    v.wl = d.wl;
    v.high = bv_view(d.high);
    v.low = d.low;
    return v;
}

static uint32_t hi_bit(uint64_t x)
{
    uint32_t r = 0;
    while (x >>= 1)
        ++r;
    return r;
}

int sd_build(sdslgpu_handle * h, uint64_t const * words_in, bool on_device, uint64_t nbits, cudaStream_t s)
{
    SdImage & d = h->sd;
    d.size = nbits;
    uint64_t nwords = (nbits + 63) >> 6;
    uint64_t * words = nullptr;
    SG_TRY(h->pool.alloc_t(&words, nwords + 2));
    SG_CUDA(cudaMemsetAsync(words + nwords, 0, 16, s));
    if (nwords)
        SG_CUDA(cudaMemcpyAsync(words, words_in, nwords * 8, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s));
    uint32_t * cnt = nullptr;
    uint64_t *before = nullptr, *tmp = nullptr;
    SG_TRY(h->pool.alloc_t(&cnt, nwords + 1));
    SG_TRY(h->pool.alloc_t(&before, nwords + 1));
    SG_TRY(h->pool.alloc_t(&tmp, scan_tmp_words(nwords)));
    if (nwords)
    {
        sd_count_kernel<<<blocks_for(nwords), kThreads, 0, s>>>(words, nbits, nwords, cnt);
        SG_CUDA(cudaGetLastError());
    }
    SG_CUDA(exclusive_scan(cnt, nwords, before, tmp, s));
    SG_CUDA(cudaMemcpyAsync(&d.m, before + nwords, 8, cudaMemcpyDeviceToHost, s));
    SG_CUDA(cudaStreamSynchronize(s));
    // sd_vector.hpp:222-228
    uint32_t logm = hi_bit(d.m) + 1, logn = hi_bit(nbits) + 1;
    if (logm == logn)
        --logm;
    d.wl = logn - logm;
    d.high_bits = d.m + (1ull << logm);
    d.low_words = ((d.m * d.wl + 63) >> 6) + 2;
    uint64_t high_words = ((d.high_bits + 63) >> 6) + 2;
    uint64_t * high = nullptr;
    SG_TRY(h->pool.alloc_t(&d.low, d.low_words));
    SG_TRY(h->pool.alloc_t(&high, high_words));
    SG_CUDA(cudaMemsetAsync(d.low, 0, d.low_words * 8, s));
    SG_CUDA(cudaMemsetAsync(high, 0, high_words * 8, s));
    if (nwords)
    {
        sd_scatter_kernel<<<blocks_for(nwords), kThreads, 0, s>>>(words, nbits, nwords, before, d.wl, reinterpret_cast<unsigned long long *>(d.low),
                                                                  reinterpret_cast<unsigned long long *>(high));
        SG_CUDA(cudaGetLastError());
    }
    SG_CUDA(cudaStreamSynchronize(s));
    h->pool.release(words);
    h->pool.release(cnt);
    h->pool.release(before);
    h->pool.release(tmp);
    // rank blocks + select samples over `high`; SDSLGPU_F_SDSL_LAYOUT keeps the raw words for serialisation
    SG_TRY(bv_build(h->pool, d.high, h->flags & SDSLGPU_F_SDSL_LAYOUT, high, true, d.high_bits, s));
    h->pool.release(high);
    return SDSLGPU_OK;
}

// ops of the locality-ordered batch pipeline (binned.cuh): rank(i) looks at bucket i >> wl of `high` and the low
// parts next to it, select_1(i) at low[i-1] and the i-th one of `high` — both positions grow with the key
struct SdRankOp
{
    static constexpr int kIlp = 1;
    static constexpr int kMinCtas = 6;
    static constexpr uint32_t kSmem = 0;
    SdView v;
    int b;
    __device__ __forceinline__ void stage(uint8_t *) const
    {}
    __device__ __forceinline__ uint64_t operator()(uint64_t i) const
    {
        uint64_t r = sd_rank1_one(v, i);
        return b ? r : i - r;
    }
};
struct SdSelect1Op
{
    static constexpr int kIlp = 1;
    static constexpr int kMinCtas = 6;
    static constexpr uint32_t kSmem = 0;
    SdView v;
    __device__ __forceinline__ void stage(uint8_t *) const
    {}
    __device__ __forceinline__ uint64_t operator()(uint64_t key) const
    {
        return sd_select1_one(v, key + 1);
    }
};

static uint64_t sd_index_bytes(SdImage const & d)
{
    return d.high.nblocks * sizeof(bvblock) + d.low_words * 8;
}

int sd_rank_device(sdslgpu_handle const * h, int b, uint64_t const * idx, uint64_t n, uint64_t * out, cudaStream_t s)
{
    if (n == 0)
        return SDSLGPU_OK;
    if (bin_wanted(h->order, sd_index_bytes(h->sd), n))
    {
        bool done = false;
        SG_TRY(bin_run(SdRankOp{sd_view(h->sd), b}, sd_index_bytes(h->sd), 0, h->sd.size, idx, n, out, s, &done));
        if (done)
            return SDSLGPU_OK;
    }
    sd_rank_kernel<<<grid_for(n), kThreads, 0, s>>>(sd_view(h->sd), b, idx, n, out);
    SG_CUDA(cudaGetLastError());
    return SDSLGPU_OK;
}

int sd_select_device(sdslgpu_handle const * h, int b, uint64_t const * idx, uint64_t n, uint64_t * out, cudaStream_t s)
{
    if (n == 0)
        return SDSLGPU_OK;
    if (b && h->sd.m && bin_wanted(h->order, sd_index_bytes(h->sd), n, kBinSelectDensity)) // select_0 is a binary search over select_1: no locality
    {
        bool done = false;
        SG_TRY(bin_run(SdSelect1Op{sd_view(h->sd)}, sd_index_bytes(h->sd), 1, h->sd.m - 1, idx, n, out, s, &done));
        if (done)
            return SDSLGPU_OK;
    }
    if (b)
        sd_select_kernel<1><<<grid_for(n), kThreads, 0, s>>>(sd_view(h->sd), idx, n, out);
    else
        sd_select_kernel<0><<<grid_for(n), kThreads, 0, s>>>(sd_view(h->sd), idx, n, out);
    SG_CUDA(cudaGetLastError());
    return SDSLGPU_OK;
}

int sd_access_device(sdslgpu_handle const * h, uint64_t const * idx, uint64_t n, uint64_t * out, cudaStream_t s)
{
    if (n == 0)
        return SDSLGPU_OK;
    sd_access_kernel<<<grid_for(n), kThreads, 0, s>>>(sd_view(h->sd), idx, n, out);
    SG_CUDA(cudaGetLastError());
    return SDSLGPU_OK;
}

// m_low and m_high exactly as the reference serialises them (sd_vector.hpp:426-433, without the two select
// supports): construction parity of the device builder
int sd_serialize_low_high(sdslgpu_handle const * h, std::vector<uint8_t> & blob)
{
    SdImage const & d = h->sd;
    if (!d.high.words)
    {
        set_error("sd serialisation needs a handle created with SDSLGPU_F_SDSL_LAYOUT");
        return SDSLGPU_ENOTSUP;
    }
    auto put64 = [&](uint64_t x) {
        for (int k = 0; k < 8; ++k)
            blob.push_back((uint8_t)(x >> (8 * k)));
    };
    put64(d.size);
    blob.push_back((uint8_t)d.wl);
    uint64_t lbits = d.m * d.wl, lw = (lbits + 63) >> 6, hw = (d.high_bits + 63) >> 6;
    std::vector<uint64_t> low(lw + 1), high(hw + 1);
    if (lw)
        SG_CUDA(cudaMemcpy(low.data(), d.low, lw * 8, cudaMemcpyDeviceToHost));
    SG_CUDA(cudaMemcpy(high.data(), d.high.words, hw * 8, cudaMemcpyDeviceToHost));
    put64(((uint64_t)d.wl << 56) | lbits);
    for (uint64_t k = 0; k < lw; ++k)
        put64(low[k]);
    put64((1ull << 56) | d.high_bits);
    for (uint64_t k = 0; k < hw; ++k)
        put64(high[k]);
    return SDSLGPU_OK;
}

} // namespace sdslgpu
