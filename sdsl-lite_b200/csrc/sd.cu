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
            {
                r = v.samp0 ? sd_select0_one(v, i) : ~0ull;
                if (r == ~0ull)
                    r = sd_select0_bsearch(v, i);
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
    v.samp0 = d.samp0;
    v.log_s0 = d.log_s0;
    return v;
}

// one thread per sector block of `high`: the samples whose zero is crossed in this block (sd_device.cuh)
__global__ void __launch_bounds__(kThreads) sd_samp0_kernel(SdView const v, uint64_t nblocks, uint32_t log_s, uint64_t nsamp, uint32_t * __restrict__ samp0)
{
    uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= nblocks)
        return;
    int64_t v_prev, v_last;
    if (!sd_block_zero_span(v, g, v_prev, v_last))
        return;
    // sample q stands for zero number q * S + 1: owned iff v_prev < q * S + 1 <= v_last
    uint64_t const S = 1ull << log_s;
    for (uint64_t q = ((uint64_t)v_prev + S - 1) >> log_s; q < nsamp && (q << log_s) + 1 <= (uint64_t)v_last; ++q)
        samp0[q] = (uint32_t)g;
}

int sd_build_select0_samples(sdslgpu_handle * h, cudaStream_t s)
{
    SdImage & d = h->sd;
    uint64_t const zeros = d.size - d.m;
    if (zeros == 0 || d.high.nblocks >= (1ull << 32) || (h->flags & SDSLGPU_F_NO_SELECT))
        return SDSLGPU_OK;
    // about one sample per sector block of `high`: the sample's block is then the crossing block or its neighbour
    uint32_t ls = 0;
    while ((zeros >> ls) > d.high.nblocks)
        ++ls;
    if (char const * e = std::getenv("SDSLGPU_SD_SELECT0_LOG_S")) // tuning / test knob
        ls = (uint32_t)std::atoi(e) > 40 ? 40u : (uint32_t)std::atoi(e);
    d.log_s0 = ls;
    d.nsamp0 = ((zeros - 1) >> ls) + 1;
    SG_TRY(h->pool.alloc_t(&d.samp0, d.nsamp0 + 1));
    SG_CUDA(cudaMemsetAsync(d.samp0, 0, (d.nsamp0 + 1) * 4, s));
    SdView v = sd_view(d);
    v.samp0 = nullptr;
    sd_samp0_kernel<<<blocks_for(d.high.nblocks), kThreads, 0, s>>>(v, d.high.nblocks, ls, d.nsamp0, d.samp0);
    SG_CUDA(cudaGetLastError());
    SG_CUDA(cudaStreamSynchronize(s));
    return SDSLGPU_OK;
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
    return sd_build_select0_samples(h, s);
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

// select_0(i): the crossing block of `high` and the low parts of its bucket both move forward with i
struct SdSelect0Op
{
    static constexpr int kIlp = 1;
    static constexpr int kMinCtas = 6;
    static constexpr uint32_t kSmem = 0;
    SdView v;
    __device__ __forceinline__ void stage(uint8_t *) const
    {}
    __device__ __forceinline__ uint64_t operator()(uint64_t key) const
    {
        uint64_t r = sd_select0_one(v, key + 1);
        return r != ~0ull ? r : sd_select0_bsearch(v, key + 1);
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
    if (b && h->sd.m && bin_wanted(h->order, sd_index_bytes(h->sd), n, kBinSelectDensity))
    {
        bool done = false;
        SG_TRY(bin_run(SdSelect1Op{sd_view(h->sd)}, sd_index_bytes(h->sd), 1, h->sd.m - 1, idx, n, out, s, &done));
        if (done)
            return SDSLGPU_OK;
    }
    if (!b && h->sd.samp0 && h->sd.size > h->sd.m && bin_wanted(h->order, sd_index_bytes(h->sd), n, kBinSelectDensity))
    { // (without samples select_0 is a binary search over select_1: no locality to order the batch by)
        bool done = false;
        SG_TRY(bin_run(SdSelect0Op{sd_view(h->sd)}, sd_index_bytes(h->sd), 1, h->sd.size - h->sd.m - 1, idx, n, out, s, &done));
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
