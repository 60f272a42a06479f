// fm.cu — csa_wt<wt_huff<>, 32, 64> (FM-index): construction and the batched search kernels.
//
// Replaces (results bit-exact):
//   construct(csa, text)      construct.hpp:127-193 + csa_wt.hpp:323-355 -> csa_build_from_text
//        (suffix array by SA-IS on the host instead of the vendored divsufsort; BWT, byte_alphabet
//         csa_alphabet_strategy.hpp:175-212, SA samples csa_sampling_strategy.hpp:98-115)
//   backward_search / count   suffix_array_algorithm.hpp:166-248, 463-471 -> fm_count_kernel
//   csa_wt::operator[]        csa_wt.hpp:363-381 (+ LF, suffix_array_helper.hpp:346-360) -> fm_sa_kernel
//   locate                    suffix_array_algorithm.hpp:534-550 -> fm_count_kernel + scan + fm_locate_fill_kernel
// One thread per pattern / per occurrence; the two rank chains of a backward-search step walk the same
// root-to-leaf path, so their sector gathers are issued together.  Node table, paths, C[] and char2comp[]
// (~19 KB) are staged in shared memory once per CTA.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <thread>

#include "internal.h"
#include "wt_device.cuh"
#include "sais.h"
#include "scan.cuh"

namespace sdslgpu
{

static constexpr uint16_t kUndef = kWtUndef;

struct FmSmem
{
    WtTree tree;
    FmTables tab;
};

__device__ __forceinline__ void stage_fm(WtTree const * __restrict__ gt, FmTables const * __restrict__ gf, FmSmem * s)
{
    uint4 const * a = reinterpret_cast<uint4 const *>(gt);
    uint4 * da = reinterpret_cast<uint4 *>(&s->tree);
    for (uint32_t k = threadIdx.x; k < sizeof(WtTree) / 16; k += blockDim.x)
        da[k] = __ldg(a + k);
    uint4 const * b = reinterpret_cast<uint4 const *>(gf);
    uint4 * db = reinterpret_cast<uint4 *>(&s->tab);
    for (uint32_t k = threadIdx.x; k < sizeof(FmTables) / 16; k += blockDim.x)
        db[k] = __ldg(b + k);
    __syncthreads();
}

// (rank(a, c), rank(b, c)) on the BWT's wavelet tree: both chains descend the same path (wt_pc.hpp:371-399)
template <class Bits>
__device__ __forceinline__ void wt_rank_pair(Bits const & bits, WtTree const * t, uint64_t sigma, uint32_t c, uint64_t & a, uint64_t & b)
{
    if (t->c_to_leaf[c] == kUndef)
    {
        a = b = 0;
        return;
    }
    if (sigma == 1)
        return;
    uint64_t p = t->path[c];
    uint32_t len = (uint32_t)(p >> 56);
    uint32_t v = 0;
    for (uint32_t l = 0; l < len && (a | b); ++l, p >>= 1)
    {
        uint64_t base = t->bv_pos[v], br = t->bv_pos_rank[v];
        uint64_t oa = bits.rank1(base + a) - br;
        uint64_t ob = bits.rank1(base + b) - br;
        a = (p & 1) ? oa : a - oa;
        b = (p & 1) ? ob : b - ob;
        v = t->child[v][p & 1];
    }
}

template <class Bits>
__global__ void __launch_bounds__(kThreads) fm_count_kernel(Bits bits,
                                                            WtTree const * __restrict__ tree,
                                                            FmTables const * __restrict__ tab,
                                                            uint64_t n,     // csa.size() = text length + 1
                                                            uint64_t sigma, // of the BWT's wavelet tree
                                                            uint8_t const * __restrict__ pats,
                                                            uint64_t const * __restrict__ off,
                                                            uint64_t npat,
                                                            uint64_t * __restrict__ cnt_out,
                                                            uint64_t * __restrict__ l_out)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FmSmem * sm = reinterpret_cast<FmSmem *>(smem_raw);
    stage_fm(tree, tab, sm);
    bits.attach(smem_raw + sizeof(FmSmem));
    uint64_t const stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < npat; q += stride)
    {
        uint64_t b = off[q], e = off[q + 1];
        uint64_t l = 0, r = n - 1;
        if (e - b > n) // count(): a pattern longer than the text cannot occur (suffix_array_algorithm.hpp:466-467)
        {
            cnt_out[q] = 0;
            if (l_out)
                l_out[q] = 0;
            continue;
        }
        uint64_t it = e;
        while (it > b && r + 1 - l > 0)
        {
            --it;
            uint32_t c = pats[it];
            uint32_t cc = sm->tab.char2comp[c];
            if (cc == 0 && c > 0)
            { // character not in the text (suffix_array_algorithm.hpp:180-184)
                l = 1;
                r = 0;
            }
            else
            {
                uint64_t cb = sm->tab.C[cc];
                if (l == 0 && r + 1 == n)
                { // full interval: table only (:188-192)
                    l = cb;
                    r = sm->tab.C[cc + 1] - 1;
                }
                else
                {
                    uint64_t ra = l, rb = r + 1;
                    wt_rank_pair(bits, &sm->tree, sigma, c, ra, rb);
                    l = cb + ra;
                    r = cb + rb - 1;
                }
            }
        }
        cnt_out[q] = r + 1 - l;
        if (l_out)
            l_out[q] = l;
    }
}

// SA[i] by LF-walking to the next sampled index (csa_wt.hpp:363-381)
template <class Bits>
__device__ __forceinline__ uint64_t fm_sa_one(Bits const & bits, FmSmem const * sm, uint64_t const * __restrict__ samples, uint32_t dens, uint64_t n, uint64_t i)
{
    uint64_t steps = 0;
    while (i % dens != 0)
    {
        uint32_t sym;
        uint64_t j = wt_inverse_select_one(bits, &sm->tree, i, sym);
        i = sm->tab.C[sm->tab.char2comp[sym]] + j; // LF (suffix_array_helper.hpp:352-359)
        ++steps;
    }
    uint64_t v = __ldg(samples + i / dens) + steps;
    return v < n ? v : v - n;
}

template <class Bits>
__global__ void __launch_bounds__(kThreads) fm_sa_kernel(Bits bits,
                                                         WtTree const * __restrict__ tree,
                                                         FmTables const * __restrict__ tab,
                                                         uint64_t const * __restrict__ samples,
                                                         uint32_t dens,
                                                         uint64_t n,
                                                         uint64_t const * __restrict__ idx,
                                                         uint64_t cnt,
                                                         uint64_t * __restrict__ out)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FmSmem * sm = reinterpret_cast<FmSmem *>(smem_raw);
    stage_fm(tree, tab, sm);
    bits.attach(smem_raw + sizeof(FmSmem));
    uint64_t const stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < cnt; q += stride)
    {
        uint64_t i = idx[q];
        out[q] = (i < n) ? fm_sa_one(bits, sm, samples, dens, n, i) : SDSLGPU_NPOS;
    }
}

// locate, phase 3: one thread per reported occurrence; occ[occ_off[k] + j] = SA[l[k] + j]  (SA order)
template <class Bits>
__global__ void __launch_bounds__(kThreads) fm_locate_fill_kernel(Bits bits,
                                                                  WtTree const * __restrict__ tree,
                                                                  FmTables const * __restrict__ tab,
                                                                  uint64_t const * __restrict__ samples,
                                                                  uint32_t dens,
                                                                  uint64_t n,
                                                                  uint64_t const * __restrict__ l,
                                                                  uint64_t const * __restrict__ occ_off,
                                                                  uint64_t npat,
                                                                  uint64_t total,
                                                                  uint64_t * __restrict__ occ)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FmSmem * sm = reinterpret_cast<FmSmem *>(smem_raw);
    stage_fm(tree, tab, sm);
    bits.attach(smem_raw + sizeof(FmSmem));
    uint64_t const stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t o = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += stride)
    {
        // largest k with occ_off[k] <= o
        uint64_t lo = 0, hi = npat - 1;
        while (lo < hi)
        {
            uint64_t mid = (lo + hi + 1) >> 1;
            if (__ldg(occ_off + mid) <= o)
                lo = mid;
            else
                hi = mid - 1;
        }
        uint64_t j = o - __ldg(occ_off + lo);
        occ[o] = fm_sa_one(bits, sm, samples, dens, n, __ldg(l + lo) + j);
    }
}

// extract(csa, begin, end) (suffix_array_algorithm.hpp:590-610): text[begin..end] by walking LF backwards from
// ISA[end]; ISA[end] itself is reached from the next ISA sample (suffix_array_helper.hpp:519-537).  One thread
// per requested range; the chain is sequential, the batch is the parallelism.
template <class Bits>
__global__ void __launch_bounds__(kThreads) fm_extract_kernel(Bits bits,
                                                              WtTree const * __restrict__ tree,
                                                              FmTables const * __restrict__ tab,
                                                              uint64_t const * __restrict__ isa_samples,
                                                              uint64_t nisa,
                                                              uint32_t isa_dens,
                                                              uint64_t n,
                                                              uint64_t const * __restrict__ begin,
                                                              uint64_t const * __restrict__ end,
                                                              uint64_t const * __restrict__ out_off,
                                                              uint64_t cnt,
                                                              uint8_t * __restrict__ out)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FmSmem * sm = reinterpret_cast<FmSmem *>(smem_raw);
    stage_fm(tree, tab, sm);
    bits.attach(smem_raw + sizeof(FmSmem));
    uint64_t const stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < cnt; q += stride)
    {
        uint64_t b = begin[q], e = end[q];
        if (e >= n || b > e)
            continue; // out of domain: the caller sized the output from (begin, end); nothing is written
        uint8_t * dst = out + out_off[q];
        // ISA[e]: rightmost... the next sample at or after e, then LF back (sample_qeq, csa_sampling_strategy.hpp:795-800)
        uint64_t ci = (e / isa_dens + 1) % nisa, pos = ci * isa_dens;
        uint64_t order = __ldg(isa_samples + ci);
        uint64_t back = pos < e ? pos + n - e : pos - e;
        while (back--)
        {
            uint32_t sym;
            uint64_t j = wt_inverse_select_one(bits, &sm->tree, order, sym);
            order = sm->tab.C[sm->tab.char2comp[sym]] + j;
        }
        uint64_t steps = e - b + 1;
        // first_row_symbol(order): the symbol whose C-bucket holds `order`
        uint32_t lo = 0, hi = sm->tab.sigma;
        while (hi - lo > 1)
        {
            uint32_t mid = (lo + hi) >> 1;
            if (sm->tab.C[mid] <= order)
                lo = mid;
            else
                hi = mid;
        }
        dst[--steps] = sm->tab.comp2char[lo];
        while (steps != 0)
        {
            uint32_t sym;
            uint64_t j = wt_inverse_select_one(bits, &sm->tree, order, sym);
            order = sm->tab.C[sm->tab.char2comp[sym]] + j;
            dst[--steps] = (uint8_t)sym;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// host: construction
// ------------------------------------------------------------------------------------------------
int csa_build_from_text(sdslgpu_handle * h, uint8_t const * text, uint64_t len, uint32_t sa_dens, uint32_t isa_dens, cudaStream_t s)
{
    if (void const * z = len ? std::memchr(text, 0, len) : nullptr)
    {
        set_error("csa: the text contains a zero byte at position %llu (the reference rejects it too, construct.hpp:34-46)",
                  (unsigned long long)(static_cast<uint8_t const *>(z) - text));
        return SDSLGPU_EINVAL;
    }
    uint64_t n = len + 1;
    CsaImage & c = h->csa;
    c.n = n;
    c.sa_dens = sa_dens ? sa_dens : 32;
    std::vector<uint8_t> bwt;
    std::vector<uint64_t> samples, isa;
    c.isa_dens = isa_dens ? isa_dens : 64;
    // suffix array + BWT + samples: on the device (prefix doubling, gpu_sa.cu) unless SDSLGPU_HOST_SA=1 or the
    // text does not fit 32-bit suffix indices / device memory, in which case the host SA-IS builder runs
    int st = SDSLGPU_ENOTSUP;
    char const * force_host = std::getenv("SDSLGPU_HOST_SA");
    if (!(force_host && std::atoi(force_host) != 0))
        st = gpu_suffix_array_bwt(text, len, c.sa_dens, c.isa_dens, bwt, samples, isa, nullptr, s);
    if (st == SDSLGPU_ENOTSUP)
    {
        std::vector<uint8_t> t(n);
        bwt.assign(n, 0);
        std::memcpy(t.data(), text, len);
        t[len] = 0;
        samples.assign((n + c.sa_dens - 1) / c.sa_dens, 0);
        isa.assign((n - 1) / c.isa_dens + 1, 0);
        if (n < (1ull << 31))
        {
            std::vector<int32_t> sa(n);
            sais<uint8_t, int32_t>(t.data(), sa.data(), (int32_t)n, 255);
            for (uint64_t i = 0; i < n; ++i)
                bwt[i] = sa[i] ? t[sa[i] - 1] : t[n - 1];
            for (uint64_t i = 0; i < n; i += c.sa_dens)
                samples[i / c.sa_dens] = (uint64_t)sa[i];
            for (uint64_t i = 0; i < n; ++i)
                if ((uint64_t)sa[i] % c.isa_dens == 0)
                    isa[(uint64_t)sa[i] / c.isa_dens] = i;
        }
        else
        {
            std::vector<int64_t> sa(n);
            sais<uint8_t, int64_t>(t.data(), sa.data(), (int64_t)n, 255);
            for (uint64_t i = 0; i < n; ++i)
                bwt[i] = sa[i] ? t[sa[i] - 1] : t[n - 1];
            for (uint64_t i = 0; i < n; i += c.sa_dens)
                samples[i / c.sa_dens] = (uint64_t)sa[i];
            for (uint64_t i = 0; i < n; ++i)
                if ((uint64_t)sa[i] % c.isa_dens == 0)
                    isa[(uint64_t)sa[i] / c.isa_dens] = i;
        }
    }
    else if (st != SDSLGPU_OK)
        return st;
    // byte_alphabet (csa_alphabet_strategy.hpp:175-212)
    FmTables & tab = c.host_tab;
    std::memset(&tab, 0, sizeof(tab));
    uint64_t cnt[256] = {0};
    {
        unsigned T = std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
        if (n < (1u << 20))
            T = 1;
        std::vector<std::vector<uint64_t>> part(T, std::vector<uint64_t>(256, 0));
        std::vector<std::thread> th;
        uint64_t chunk = (n + T - 1) / T;
        for (unsigned t = 0; t < T; ++t)
            th.emplace_back([&, t] {
                uint64_t lo = std::min(n, t * chunk), hi = std::min(n, lo + chunk);
                uint64_t * c = part[t].data();
                for (uint64_t k = lo; k < hi; ++k)
                    ++c[bwt[k]];
            });
        for (auto & x : th)
            x.join();
        for (unsigned t = 0; t < T; ++t)
            for (int ch = 0; ch < 256; ++ch)
                cnt[ch] += part[t][ch];
    }
    uint32_t sigma = 0;
    for (int ch = 0; ch < 256; ++ch)
        if (cnt[ch])
        {
            tab.char2comp[ch] = (uint8_t)sigma;
            tab.comp2char[sigma] = (uint8_t)ch;
            tab.C[sigma + 1] = cnt[ch];
            ++sigma;
        }
    for (uint32_t k = 1; k <= sigma; ++k)
        tab.C[k] += tab.C[k - 1];
    tab.sigma = sigma;
    return csa_upload(h, bwt.data(), samples.data(), samples.size(), isa.data(), isa.size(), s);
}

// uploads the CSA parts; the wavelet tree of the BWT is built by the wt_huff path
int csa_upload_isa(sdslgpu_handle * h, uint64_t const * isa, uint64_t nisa, cudaStream_t s)
{
    CsaImage & c = h->csa;
    c.nisa = nisa;
    SG_TRY(h->pool.alloc_t(&c.isa_samples, nisa + 1));
    SG_CUDA(cudaMemcpyAsync(c.isa_samples, isa, nisa * 8, cudaMemcpyHostToDevice, s));
    SG_CUDA(cudaStreamSynchronize(s));
    return SDSLGPU_OK;
}

int csa_upload(sdslgpu_handle * h, uint8_t const * bwt, uint64_t const * samples, uint64_t nsamples, uint64_t const * isa, uint64_t nisa, cudaStream_t s)
{
    CsaImage & c = h->csa;
    SG_TRY(wt_huff_build_from_text(h, bwt, c.n, s));
    SG_TRY(csa_upload_isa(h, isa, nisa, s));
    c.nsamples = nsamples;
    SG_TRY(h->pool.alloc_t(&c.samples, nsamples + 1));
    SG_CUDA(cudaMemcpyAsync(c.samples, samples, nsamples * 8, cudaMemcpyHostToDevice, s));
    SG_TRY(h->pool.alloc_t(&c.tab, 1));
    SG_CUDA(cudaMemcpyAsync(c.tab, &c.host_tab, sizeof(FmTables), cudaMemcpyHostToDevice, s));
    SG_CUDA(cudaStreamSynchronize(s));
    // the one-hot occurrence bitmaps the searches run on (occ16_device.cuh), unless the caller asked for a compact index
    if (!(h->flags & (SDSLGPU_F_COMPACT | SDSLGPU_F_RRR_BV)))
        SG_TRY(occ16_build(h, bwt, s));
    return SDSLGPU_OK;
}

// ------------------------------------------------------------------------------------------------
// launchers (all pointers are device pointers here)
// ------------------------------------------------------------------------------------------------
static size_t const kFmSmem = sizeof(FmSmem);

int fm_count_device(sdslgpu_handle const * h, uint8_t const * pats, uint64_t const * off, uint64_t npat, uint64_t * cnt, uint64_t * l, cudaStream_t s)
{
    if (h->csa.occ.levels)
        return fm16_count_device(h, pats, off, npat, cnt, l, s);
    if (npat == 0)
        return SDSLGPU_OK;
    SG_LAUNCH_BITS(fm_count_kernel, h->wt, grid_for(npat), kFmSmem, s, h->wt.tree, h->csa.tab, h->csa.n, h->wt.sigma, pats, off, npat, cnt, l);
    SG_CUDA(cudaGetLastError());
    return SDSLGPU_OK;
}

int fm_sa_device(sdslgpu_handle const * h, uint64_t const * idx, uint64_t cnt, uint64_t * out, cudaStream_t s)
{
    if (h->csa.occ.levels)
        return fm16_sa_device(h, idx, cnt, out, s);
    if (cnt == 0)
        return SDSLGPU_OK;
    SG_LAUNCH_BITS(fm_sa_kernel, h->wt, grid_for(cnt), kFmSmem, s, h->wt.tree, h->csa.tab, h->csa.samples, h->csa.sa_dens, h->csa.n, idx, cnt, out);
    SG_CUDA(cudaGetLastError());
    return SDSLGPU_OK;
}

int fm_extract_device(sdslgpu_handle const * h, uint64_t const * begin, uint64_t const * end, uint64_t const * out_off, uint64_t n, uint8_t * out, cudaStream_t s)
{
    if (h->csa.occ.levels)
        return fm16_extract_device(h, begin, end, out_off, n, out, s);
    if (n == 0)
        return SDSLGPU_OK;
    SG_LAUNCH_BITS(fm_extract_kernel, h->wt, grid_for(n), kFmSmem, s, h->wt.tree, h->csa.tab, h->csa.isa_samples, h->csa.nisa, h->csa.isa_dens, h->csa.n, begin, end,
                                                              out_off, n, out);
    SG_CUDA(cudaGetLastError());
    return SDSLGPU_OK;
}

// occ_off[0..npat] = exclusive prefix sums of cnt (device); tmp must hold scan_tmp_words(npat) u64
int fm_scan_counts_device(uint64_t const * cnt, uint64_t npat, uint64_t * occ_off, uint64_t * tmp, cudaStream_t s)
{
    SG_CUDA(exclusive_scan<uint64_t>(cnt, npat, occ_off, tmp, s));
    return SDSLGPU_OK;
}

int fm_locate_fill_device(sdslgpu_handle const * h, uint64_t const * l, uint64_t const * occ_off, uint64_t npat, uint64_t total, uint64_t * occ, cudaStream_t s)
{
    if (h->csa.occ.levels)
        return fm16_locate_fill_device(h, l, occ_off, npat, total, occ, s);
    if (total == 0 || npat == 0)
        return SDSLGPU_OK;
    SG_LAUNCH_BITS(fm_locate_fill_kernel, h->wt, grid_for(total), kFmSmem, s, h->wt.tree, h->csa.tab, h->csa.samples, h->csa.sa_dens, h->csa.n, l, occ_off, npat,
                                                                      total, occ);
    SG_CUDA(cudaGetLastError());
    return SDSLGPU_OK;
}

} // namespace sdslgpu

namespace sdslgpu
{
uint64_t fm_scan_tmp_words(uint64_t npat)
{
    return scan_tmp_words(npat);
}
} // namespace sdslgpu
