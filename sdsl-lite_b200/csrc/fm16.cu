// fm16.cu — count / csa[i] / locate / extract over the one-hot occurrence structure (occ16_device.cuh), and
// its construction on the device.  The kernels are the ones of fm.cu with the two wavelet-tree primitives
// swapped: backward_search step (suffix_array_algorithm.hpp:166-248) = 2 gathers for both interval ends when
// the interval is narrow, LF (suffix_array_helper.hpp:346-360) = 1 byte + 2 gathers.  Results are bit-identical
// to the wavelet-tree path (tests/test_fm_gpu.py runs every case on both).
#include <cstring>
#include <vector>

#include "internal.h"
#include "occ16_device.cuh"
#include "wt_device.cuh"

namespace sdslgpu
{

struct Fm16Smem
{
    FmTables tab;
    Occ16Tab occ;
};

__device__ __forceinline__ void stage_fm16(FmTables const * __restrict__ gf, Occ16Tab const * __restrict__ go, Fm16Smem * s)
{
    uint4 const * a = reinterpret_cast<uint4 const *>(gf);
    uint4 * da = reinterpret_cast<uint4 *>(&s->tab);
    for (uint32_t k = threadIdx.x; k < sizeof(FmTables) / 16; k += blockDim.x)
        da[k] = __ldg(a + k);
    uint4 const * b = reinterpret_cast<uint4 const *>(go);
    uint4 * db = reinterpret_cast<uint4 *>(&s->occ);
    for (uint32_t k = threadIdx.x; k < sizeof(Occ16Tab) / 16; k += blockDim.x)
        db[k] = __ldg(b + k);
    __syncthreads();
}

__global__ void __launch_bounds__(kThreads) fm16_count_kernel(FmTables const * __restrict__ tab,
                                                              Occ16Tab const * __restrict__ occ,
                                                              uint64_t n,
                                                              uint8_t const * __restrict__ pats,
                                                              uint64_t const * __restrict__ off,
                                                              uint64_t npat,
                                                              uint64_t * __restrict__ cnt_out,
                                                              uint64_t * __restrict__ l_out)
{
    __shared__ Fm16Smem sm;
    stage_fm16(tab, occ, &sm);
    uint64_t const stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < npat; q += stride)
    {
        uint64_t b = off[q], e = off[q + 1];
        uint64_t lo = 0, hi = n; // rows [lo, hi)
        if (e - b > n)
            hi = 0; // a pattern longer than the text cannot occur (suffix_array_algorithm.hpp:466-467)
        uint64_t it = e;
        while (it > b && hi > lo)
        {
            --it;
            uint32_t c = pats[it];
            uint32_t cc = sm.tab.char2comp[c];
            if (cc == 0 && c > 0)
            { // character not in the text (:180-184)
                lo = 1;
                hi = 1;
            }
            else if (lo == 0 && hi == n)
            { // full interval: table only (:188-192)
                lo = sm.tab.C[cc];
                hi = sm.tab.C[cc + 1];
            }
            else
                occ16_backward_step(&sm.occ, cc, lo, hi);
        }
        cnt_out[q] = hi - lo;
        if (l_out)
            l_out[q] = (e - b > n) ? 0 : lo;
    }
}

// SA[i] by LF-walking to the next sampled row (csa_wt.hpp:363-381)
__device__ __forceinline__ uint64_t fm16_sa_one(Fm16Smem const * sm, uint64_t const * __restrict__ samples, uint32_t dens, uint64_t n, uint64_t i)
{
    uint64_t steps = 0;
    while (i % dens != 0)
    {
        uint32_t cc;
        i = occ16_lf(&sm->occ, i, cc);
        ++steps;
    }
    uint64_t v = __ldg(samples + i / dens) + steps;
    return v < n ? v : v - n;
}

__global__ void __launch_bounds__(kThreads) fm16_sa_kernel(FmTables const * __restrict__ tab,
                                                           Occ16Tab const * __restrict__ occ,
                                                           uint64_t const * __restrict__ samples,
                                                           uint32_t dens,
                                                           uint64_t n,
                                                           uint64_t const * __restrict__ idx,
                                                           uint64_t cnt,
                                                           uint64_t * __restrict__ out)
{
    __shared__ Fm16Smem sm;
    stage_fm16(tab, occ, &sm);
    uint64_t const stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < cnt; q += stride)
    {
        uint64_t i = idx[q];
        out[q] = (i < n) ? fm16_sa_one(&sm, samples, dens, n, i) : SDSLGPU_NPOS;
    }
}

__global__ void __launch_bounds__(kThreads) fm16_locate_fill_kernel(FmTables const * __restrict__ tab,
                                                                    Occ16Tab const * __restrict__ occ,
                                                                    uint64_t const * __restrict__ samples,
                                                                    uint32_t dens,
                                                                    uint64_t n,
                                                                    uint64_t const * __restrict__ l,
                                                                    uint64_t const * __restrict__ occ_off,
                                                                    uint64_t npat,
                                                                    uint64_t total,
                                                                    uint64_t * __restrict__ out)
{
    __shared__ Fm16Smem sm;
    stage_fm16(tab, occ, &sm);
    uint64_t const stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t o = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += stride)
    {
        uint64_t lo = 0, hi = npat - 1; // largest k with occ_off[k] <= o
        while (lo < hi)
        {
            uint64_t mid = (lo + hi + 1) >> 1;
            if (__ldg(occ_off + mid) <= o)
                lo = mid;
            else
                hi = mid - 1;
        }
        out[o] = fm16_sa_one(&sm, samples, dens, n, __ldg(l + lo) + (o - __ldg(occ_off + lo)));
    }
}

// extract (suffix_array_algorithm.hpp:590-610): ISA[end] from the next ISA sample, then LF backwards
__global__ void __launch_bounds__(kThreads) fm16_extract_kernel(FmTables const * __restrict__ tab,
                                                                Occ16Tab const * __restrict__ occ,
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
    __shared__ Fm16Smem sm;
    stage_fm16(tab, occ, &sm);
    uint64_t const stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < cnt; q += stride)
    {
        uint64_t b = begin[q], e = end[q];
        if (e >= n || b > e)
            continue;
        uint8_t * dst = out + out_off[q];
        uint64_t ci = (e / isa_dens + 1) % nisa, pos = ci * isa_dens;
        uint64_t order = __ldg(isa_samples + ci);
        uint64_t back = pos < e ? pos + n - e : pos - e;
        uint32_t cc;
        while (back--)
            order = occ16_lf(&sm.occ, order, cc);
        uint64_t steps = e - b + 1;
        uint32_t lo = 0, hi = sm.tab.sigma; // first_row_symbol(order)
        while (hi - lo > 1)
        {
            uint32_t mid = (lo + hi) >> 1;
            if (sm.tab.C[mid] <= order)
                lo = mid;
            else
                hi = mid;
        }
        dst[--steps] = sm.tab.comp2char[lo];
        while (steps != 0)
        {
            order = occ16_lf(&sm.occ, order, cc);
            dst[--steps] = sm.tab.comp2char[cc];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// construction
// ------------------------------------------------------------------------------------------------
__global__ void occ16_comp_kernel(uint8_t * __restrict__ sym, uint64_t n, FmTables const * __restrict__ tab)
{
    __shared__ uint8_t map[256];
    map[threadIdx.x] = tab->char2comp[threadIdx.x];
    __syncthreads();
    uint64_t const stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        sym[i] = map[sym[i]];
}

// words[w] bit k = [((sym[64 w + k] >> shift) & 15) == v]
__global__ void occ16_words_kernel(uint8_t const * __restrict__ sym, uint64_t n, uint32_t shift, uint32_t v, uint64_t * __restrict__ words, uint64_t nwords)
{
    uint64_t const stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < nwords; w += stride)
    {
        uint64_t base = w * 64, bits = 0;
        if (base + 64 <= n)
        {
            uint4 const * p = reinterpret_cast<uint4 const *>(sym + base);
#pragma unroll
            for (uint32_t k = 0; k < 4; ++k)
            {
                uint4 x = __ldg(p + k);
                uint32_t const u[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                for (uint32_t j = 0; j < 16; ++j)
                    bits |= (uint64_t)((((u[j >> 2] >> (8 * (j & 3))) >> shift) & 15u) == v) << (16 * k + j);
            }
        }
        else
            for (uint64_t k = 0; base + k < n; ++k)
                bits |= (uint64_t)(((uint32_t)(sym[base + k] >> shift) & 15u) == v) << k;
        words[w] = bits;
    }
}

struct Occ16Level0
{
    bvblock const * blocks[16];
    uint64_t const * top[16];
    uint64_t CH[16];
};

// S[CH[h] + rank_h(i)] = sym[i] & 15: the level-1 sequence (symbols stably sorted by high nibble)
__global__ void occ16_level1_kernel(uint8_t const * __restrict__ sym, uint64_t n, Occ16Level0 lv, uint8_t * __restrict__ s1)
{
    uint64_t const stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    {
        uint32_t cc = sym[i], h = cc >> 4;
        s1[lv.CH[h] + occ16_rank1(lv.blocks[h], lv.top[h], i)] = (uint8_t)(cc & 15u);
    }
}

// symbols of the wavelet tree (ingest path: the serialised index holds the tree, not the BWT)
template <class Bits>
__global__ void __launch_bounds__(kThreads) occ16_from_wt_kernel(Bits bits, WtTree const * __restrict__ tree, uint64_t n, uint8_t * __restrict__ sym)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WtTree * t = reinterpret_cast<WtTree *>(smem_raw);
    stage_tree(tree, t);
    bits.attach(smem_raw + sizeof(WtTree));
    uint64_t const stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    {
        uint32_t c;
        wt_inverse_select_one(bits, t, i, c);
        sym[i] = (uint8_t)c;
    }
}

static int occ16_finish(sdslgpu_handle * h, uint8_t * d_sym, cudaStream_t s)
{
    CsaImage & c = h->csa;
    Occ16Image & o = c.occ;
    FmTables const & tab = c.host_tab;
    uint64_t const n = c.n;
    uint32_t const sigma = tab.sigma;
    uint32_t const levels = sigma <= 16 ? 1 : 2;
    std::memset(&o.host_tab, 0, sizeof(o.host_tab));
    o.bwtc = d_sym;
    occ16_comp_kernel<<<blocks_for(n), 256, 0, s>>>(d_sym, n, c.tab);
    SG_CUDA(cudaGetLastError());

    uint64_t const nwords = (n + 63) / 64;
    uint64_t * d_words = nullptr;
    SG_TRY(h->pool.alloc_t(&d_words, nwords + 1));
    uint32_t const shift0 = levels == 2 ? 4 : 0;
    uint32_t const nv0 = levels == 2 ? ((sigma - 1) >> 4) + 1 : sigma;
    for (uint32_t v = 0; v < nv0; ++v)
    {
        occ16_words_kernel<<<blocks_for(nwords), kThreads, 0, s>>>(d_sym, n, shift0, v, d_words, nwords);
        SG_CUDA(cudaGetLastError());
        SG_TRY(bv_build(h->pool, o.bm[0][v], SDSLGPU_F_NO_SELECT, d_words, true, n, s));
        o.host_tab.blocks[0][v] = o.bm[0][v].blocks;
        o.host_tab.top[0][v] = o.bm[0][v].top;
    }
    uint64_t cnt[256] = {0};
    for (uint32_t k = 0; k < sigma; ++k)
        cnt[k] = tab.C[k + 1] - tab.C[k];
    if (levels == 2)
    {
        Occ16Level0 lv;
        std::memset(&lv, 0, sizeof(lv));
        uint64_t run = 0;
        for (uint32_t hh = 0; hh < 16; ++hh)
        {
            lv.blocks[hh] = o.host_tab.blocks[0][hh];
            lv.top[hh] = o.host_tab.top[0][hh];
            lv.CH[hh] = o.host_tab.CH[hh] = run;
            for (uint32_t lo = 0; lo < 16; ++lo)
                run += cnt[hh * 16 + lo];
        }
        uint8_t * d_s1 = nullptr;
        SG_TRY(h->pool.alloc_t(&d_s1, n + 64));
        SG_CUDA(cudaMemsetAsync(d_s1 + n, 0, 64, s));
        occ16_level1_kernel<<<blocks_for(n), kThreads, 0, s>>>(d_sym, n, lv, d_s1);
        SG_CUDA(cudaGetLastError());
        for (uint32_t v = 0; v < 16; ++v)
        {
            occ16_words_kernel<<<blocks_for(nwords), kThreads, 0, s>>>(d_s1, n, 0, v, d_words, nwords);
            SG_CUDA(cudaGetLastError());
            SG_TRY(bv_build(h->pool, o.bm[1][v], SDSLGPU_F_NO_SELECT, d_words, true, n, s));
            o.host_tab.blocks[1][v] = o.bm[1][v].blocks;
            o.host_tab.top[1][v] = o.bm[1][v].top;
        }
        SG_CUDA(cudaStreamSynchronize(s));
        h->pool.release(d_s1);
        // D[cc] = C[cc] - #(symbols with a smaller high nibble and the same low nibble)
        for (uint32_t k = 0; k < sigma; ++k)
        {
            uint64_t before = 0;
            for (uint32_t hh = 0; hh < (k >> 4); ++hh)
                before += cnt[hh * 16 + (k & 15u)];
            o.host_tab.D[k] = tab.C[k] - before;
        }
    }
    else
        for (uint32_t k = 0; k < sigma; ++k)
            o.host_tab.D[k] = tab.C[k];
    SG_CUDA(cudaStreamSynchronize(s));
    h->pool.release(d_words);
    o.host_tab.bwtc = d_sym;
    o.host_tab.levels = levels;
    SG_TRY(h->pool.alloc_t(&o.tab, 1));
    SG_CUDA(cudaMemcpyAsync(o.tab, &o.host_tab, sizeof(Occ16Tab), cudaMemcpyHostToDevice, s));
    SG_CUDA(cudaStreamSynchronize(s));
    o.levels = levels;
    return SDSLGPU_OK;
}

int occ16_build(sdslgpu_handle * h, uint8_t const * bwt_host, cudaStream_t s)
{
    uint64_t const n = h->csa.n;
    uint8_t * d_sym = nullptr;
    SG_TRY(h->pool.alloc_t(&d_sym, n + 64));
    SG_CUDA(cudaMemsetAsync(d_sym + n, 0, 64, s));
    SG_CUDA(cudaMemcpyAsync(d_sym, bwt_host, n, cudaMemcpyHostToDevice, s));
    return occ16_finish(h, d_sym, s);
}

int occ16_build_from_wt(sdslgpu_handle * h, cudaStream_t s)
{
    uint64_t const n = h->csa.n;
    uint8_t * d_sym = nullptr;
    SG_TRY(h->pool.alloc_t(&d_sym, n + 64));
    SG_CUDA(cudaMemsetAsync(d_sym + n, 0, 64, s));
    SG_LAUNCH_BITS(occ16_from_wt_kernel, h->wt, blocks_for(n), sizeof(WtTree), s, h->wt.tree, n, d_sym);
    SG_CUDA(cudaGetLastError());
    return occ16_finish(h, d_sym, s);
}

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
int fm16_count_device(sdslgpu_handle const * h, uint8_t const * pats, uint64_t const * off, uint64_t npat, uint64_t * cnt, uint64_t * l, cudaStream_t s)
{
    if (npat == 0)
        return SDSLGPU_OK;
    fm16_count_kernel<<<grid_for(npat), kThreads, 0, s>>>(h->csa.tab, h->csa.occ.tab, h->csa.n, pats, off, npat, cnt, l);
    SG_CUDA(cudaGetLastError());
    return SDSLGPU_OK;
}

int fm16_sa_device(sdslgpu_handle const * h, uint64_t const * idx, uint64_t cnt, uint64_t * out, cudaStream_t s)
{
    if (cnt == 0)
        return SDSLGPU_OK;
    fm16_sa_kernel<<<grid_for(cnt), kThreads, 0, s>>>(h->csa.tab, h->csa.occ.tab, h->csa.samples, h->csa.sa_dens, h->csa.n, idx, cnt, out);
    SG_CUDA(cudaGetLastError());
    return SDSLGPU_OK;
}

int fm16_locate_fill_device(sdslgpu_handle const * h, uint64_t const * l, uint64_t const * occ_off, uint64_t npat, uint64_t total, uint64_t * occ, cudaStream_t s)
{
    if (total == 0 || npat == 0)
        return SDSLGPU_OK;
    fm16_locate_fill_kernel<<<grid_for(total), kThreads, 0, s>>>(h->csa.tab, h->csa.occ.tab, h->csa.samples, h->csa.sa_dens, h->csa.n, l, occ_off, npat, total, occ);
    SG_CUDA(cudaGetLastError());
    return SDSLGPU_OK;
}

int fm16_extract_device(sdslgpu_handle const * h, uint64_t const * begin, uint64_t const * end, uint64_t const * out_off, uint64_t n, uint8_t * out, cudaStream_t s)
{
    if (n == 0)
        return SDSLGPU_OK;
    fm16_extract_kernel<<<grid_for(n), kThreads, 0, s>>>(h->csa.tab, h->csa.occ.tab, h->csa.isa_samples, h->csa.nisa, h->csa.isa_dens, h->csa.n, begin, end, out_off, n, out);
    SG_CUDA(cudaGetLastError());
    return SDSLGPU_OK;
}

} // namespace sdslgpu
