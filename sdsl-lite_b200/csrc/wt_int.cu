// wt_int.cu — wt_int<> (balanced level-wise wavelet tree over integers): construction and batched queries.
//
// Replaces (results bit-exact; the level bit vector m_tree is bit-identical to the reference's):
//   wt_int ctor               wt_int.hpp:160-260 -> wt_int_build (host stable partition per level)
//   wt_int::rank(i,c)         wt_int.hpp:379-409 -> wt_int_rank_kernel     (3 sector gathers per level)
//   wt_int::operator[] / inverse_select  :340-367 / :418-445 -> wt_int_access_kernel
//   wt_int::select(i,c)       wt_int.hpp:456-507 -> wt_int_select_kernel   (top-down path, bottom-up selects)
// The tree keeps no node table: like the reference, the node [offset, offset + node_size) is re-derived at
// every level from three rank1 calls on the concatenated level bit vector; the three gathers are independent
// and issued together.
#include <algorithm>

#include <cstdlib>

#include "internal.h"

namespace sdslgpu
{

struct WtIntView
{
    BvView tree;
    uint64_t size;
    uint32_t max_level;
};

__device__ __forceinline__ bool wt_int_has(WtIntView const & w, uint64_t c)
{
    return w.max_level >= 64 || (c >> w.max_level) == 0;
}

__global__ void __launch_bounds__(kThreads)
    wt_int_rank_kernel(WtIntView const w, uint64_t const * __restrict__ qi, uint64_t const * __restrict__ qc, uint64_t n, uint64_t * __restrict__ out)
{
    uint64_t const stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += stride)
    {
        uint64_t i = ld_stream_u64(qi + q), c = ld_stream_u64(qc + q);
        uint64_t r = SDSLGPU_NPOS;
        if (i <= w.size)
        {
            if (w.size == 0 || !wt_int_has(w, c))
                r = 0; // c is larger than any symbol (wt_int.hpp:382-385)
            else
            {
                uint64_t offset = 0, node_size = w.size, mask = 1ull << (w.max_level - 1);
                for (uint32_t k = 0; k < w.max_level && i; ++k)
                {
                    uint64_t o0 = bv_rank1(w.tree, offset);
                    uint64_t oi = bv_rank1(w.tree, offset + i) - o0;
                    uint64_t oe = bv_rank1(w.tree, offset + node_size) - o0;
                    if (c & mask)
                    {
                        offset += node_size - oe;
                        node_size = oe;
                        i = oi;
                    }
                    else
                    {
                        node_size -= oe;
                        i -= oi;
                    }
                    offset += w.size;
                    mask >>= 1;
                }
                r = i;
            }
        }
        st_stream_u64(out + q, r);
    }
}

__global__ void __launch_bounds__(kThreads)
    wt_int_access_kernel(WtIntView const w, uint64_t const * __restrict__ qi, uint64_t n, uint64_t * __restrict__ sym_out, uint64_t * __restrict__ rank_out)
{
    uint64_t const stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += stride)
    {
        uint64_t i = ld_stream_u64(qi + q);
        uint64_t c = SDSLGPU_NPOS, r = SDSLGPU_NPOS;
        if (i < w.size)
        {
            uint64_t offset = 0, node_size = w.size;
            c = 0;
            for (uint32_t k = 0; k < w.max_level; ++k)
            {
                uint32_t bit;
                uint64_t o0 = bv_rank1(w.tree, offset);
                uint64_t oi = bv_rank1_and_bit(w.tree, offset + i, bit) - o0;
                uint64_t oe = bv_rank1(w.tree, offset + node_size) - o0;
                c <<= 1;
                if (bit)
                {
                    offset += node_size - oe;
                    node_size = oe;
                    i = oi;
                    c |= 1;
                }
                else
                {
                    node_size -= oe;
                    i -= oi;
                }
                offset += w.size;
            }
            r = i;
        }
        st_stream_u64(sym_out + q, c);
        if (rank_out)
            st_stream_u64(rank_out + q, r);
    }
}

__global__ void __launch_bounds__(kThreads)
    wt_int_select_kernel(WtIntView const w, uint64_t const * __restrict__ qi, uint64_t const * __restrict__ qc, uint64_t n, uint64_t * __restrict__ out)
{
    uint64_t const stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += stride)
    {
        uint64_t i = ld_stream_u64(qi + q), c = ld_stream_u64(qc + q);
        uint64_t r = SDSLGPU_NPOS; // also the answer where the reference throws (c does not occur i times, :485-490)
        if (i >= 1 && w.size && wt_int_has(w, c))
        {
            // top-down: offsets of the nodes on c's path; only the CURRENT level's values are needed on the
            // way back up because (offset, ones before it) of level k-1 are re-derived from level k's:
            // keep them in two small local arrays like the reference (max_level <= 64)
            uint64_t path_off[65], path_rank[65];
            uint64_t offset = 0, node_size = w.size, mask = 1ull << (w.max_level - 1);
            path_off[0] = 0;
            uint32_t k = 0;
            for (; k < w.max_level && node_size; ++k)
            {
                uint64_t o0 = bv_rank1(w.tree, offset);
                uint64_t oe = bv_rank1(w.tree, offset + node_size) - o0;
                path_rank[k] = o0;
                if (c & mask)
                {
                    offset += node_size - oe;
                    node_size = oe;
                }
                else
                    node_size -= oe;
                offset += w.size;
                path_off[k + 1] = offset;
                mask >>= 1;
            }
            if (node_size != 0 && node_size >= i)
            {
                mask = 1;
                for (uint32_t lvl = w.max_level; lvl > 0; --lvl)
                {
                    uint64_t off = path_off[lvl - 1], o0 = path_rank[lvl - 1];
                    if (c & mask)
                        i = bv_select<1>(w.tree, o0 + i) - off + 1;
                    else
                        i = bv_select<0>(w.tree, off - o0 + i) - off + 1;
                    mask <<= 1;
                }
                r = i - 1;
            }
        }
        st_stream_u64(out + q, r);
    }
}

// ------------------------------------------------------------------------------------------------
// host builder: level k = the sequence stably sorted by its top k bits, bit (max_level - k - 1) of each element
// ------------------------------------------------------------------------------------------------
// Device path (wt_build.cu): sequence -> HBM, one stable radix pass per level, rank blocks + select samples from the
// device-resident level bits.  *done = false: no device memory for the scratch or SDSLGPU_HOST_WT=1 (host fill below).
static int wt_int_build_on_device(sdslgpu_handle * h, uint64_t const * seq, uint64_t n, cudaStream_t s, bool * done)
{
    *done = false;
    char const * force_host = std::getenv("SDSLGPU_HOST_WT");
    if (n == 0 || (force_host && std::atoi(force_host) != 0))
        return SDSLGPU_OK;
    struct Free
    {
        void * p;
        ~Free()
        {
            if (p)
                cudaFree(p);
        }
    };
    uint64_t * d_seq = nullptr;
    if (cudaMalloc(reinterpret_cast<void **>(&d_seq), n * 8) != cudaSuccess)
    {
        cudaGetLastError();
        return SDSLGPU_OK;
    }
    Free free_seq{d_seq};
    SG_CUDA(cudaMemcpyAsync(d_seq, seq, n * 8, cudaMemcpyHostToDevice, s));
    WtIntImage & w = h->wti;
    uint64_t * d_words = nullptr;
    uint32_t levels = 0;
    uint64_t sigma = 0;
    int st = wt_int_planes_device(d_seq, n, &levels, &sigma, &d_words, s);
    if (st == SDSLGPU_ENOTSUP)
        return SDSLGPU_OK;
    SG_TRY(st);
    Free free_words{d_words};
    w.size = n;
    w.sigma = sigma;
    w.max_level = levels;
    SG_TRY(bv_build(h->pool, w.tree, h->flags & ~SDSLGPU_F_NO_SELECT, d_words, true, n * levels, s));
    *done = true;
    return SDSLGPU_OK;
}

int wt_int_build(sdslgpu_handle * h, uint64_t const * seq, uint64_t n, cudaStream_t s)
{
    bool done = false;
    SG_TRY(wt_int_build_on_device(h, seq, n, s, &done));
    if (done)
        return SDSLGPU_OK;
    WtIntImage & w = h->wti;
    w.size = n;
    w.sigma = 0;
    w.max_level = 0;
    std::vector<uint64_t> tree(1, 0);
    uint64_t bits = 0;
    if (n)
    {
        uint64_t max_elem = 1;
        for (uint64_t i = 0; i < n; ++i)
            max_elem = std::max(max_elem, seq[i]);
        uint32_t hi = 0;
        for (uint64_t x = max_elem; x >>= 1;)
            ++hi;
        w.max_level = hi + 1;
        bits = n * w.max_level;
        tree.assign(((bits + 63) >> 6) + 1, 0);
        std::vector<uint64_t> cur(seq, seq + n), nxt(n);
        for (uint32_t k = 0; k < w.max_level; ++k)
        {
            uint32_t shift = w.max_level - k - 1;
            auto node_of = [&](uint64_t x) { return shift + 1 >= 64 ? 0ull : (x >> (shift + 1)); };
            uint64_t start = 0;
            while (start < n)
            {
                uint64_t node = node_of(cur[start]), end = start, c0 = 0;
                while (end < n && node_of(cur[end]) == node)
                {
                    c0 += !((cur[end] >> shift) & 1);
                    ++end;
                }
                uint64_t z = start, o = start + c0;
                for (uint64_t i = start; i < end; ++i)
                {
                    if ((cur[i] >> shift) & 1)
                    {
                        uint64_t pos = (uint64_t)k * n + i;
                        tree[pos >> 6] |= 1ull << (pos & 63);
                        nxt[o++] = cur[i];
                    }
                    else
                        nxt[z++] = cur[i];
                }
                if (k + 1 == w.max_level)
                    w.sigma += (c0 > 0) + (end - start - c0 > 0);
                start = end;
            }
            cur.swap(nxt);
        }
    }
    return bv_build(h->pool, w.tree, h->flags & ~SDSLGPU_F_NO_SELECT, tree.data(), false, bits, s);
}

static WtIntView wti_view(WtIntImage const & w)
{
    WtIntView v;
    v.tree = bv_view(w.tree);
    v.size = w.size;
    v.max_level = w.max_level;
    return v;
}

int wt_int_rank_device(sdslgpu_handle const * h, uint64_t const * i, uint64_t const * c, uint64_t n, uint64_t * out, cudaStream_t s)
{
    if (n == 0)
        return SDSLGPU_OK;
    wt_int_rank_kernel<<<grid_for(n), kThreads, 0, s>>>(wti_view(h->wti), i, c, n, out);
    SG_CUDA(cudaGetLastError());
    return SDSLGPU_OK;
}

int wt_int_select_device(sdslgpu_handle const * h, uint64_t const * i, uint64_t const * c, uint64_t n, uint64_t * out, cudaStream_t s)
{
    if (n == 0)
        return SDSLGPU_OK;
    wt_int_select_kernel<<<grid_for(n), kThreads, 0, s>>>(wti_view(h->wti), i, c, n, out);
    SG_CUDA(cudaGetLastError());
    return SDSLGPU_OK;
}

int wt_int_access_device(sdslgpu_handle const * h, uint64_t const * i, uint64_t n, uint64_t * sym, uint64_t * rnk, cudaStream_t s)
{
    if (n == 0)
        return SDSLGPU_OK;
    wt_int_access_kernel<<<grid_for(n), kThreads, 0, s>>>(wti_view(h->wti), i, n, sym, rnk);
    SG_CUDA(cudaGetLastError());
    return SDSLGPU_OK;
}

} // namespace sdslgpu
