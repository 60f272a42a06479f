// wt.cu — wt_huff<> (Huffman-shaped byte wavelet tree): construction and batched queries.
//
// Replaces (results bit-exact; the concatenated bit vector m_bv and the node table are built to be
// byte-identical to the reference's so that SDSL-serialised trees can be ingested as they are):
//   wt_pc ctor              wt_pc.hpp:194-248 + wt_huff.hpp:82-115 + wt_helper.hpp:230-327 -> wt_huff_build_from_text
//   wt_pc::rank(i,c)        wt_pc.hpp:371-399   -> wt_rank_kernel      (path_len dependent sector gathers)
//   wt_pc::operator[] / inverse_select  :336-357 / :411-430 -> wt_access_kernel
//   wt_pc::select(i,c)      wt_pc.hpp:443-474   -> wt_select_kernel    (bottom-up select0/select1 on m_bv)
// The node table and the per-symbol paths (≈14 KB) are staged into shared memory once per CTA.
#include <algorithm>
#include <cstdlib>
#include <atomic>
#include <queue>
#include <thread>

#include "internal.h"
#include "wt_device.cuh"
#include "wt_shape.h"

namespace sdslgpu
{

static constexpr uint16_t kUndef = kWtUndef;

// ------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------
template <class Bits>
__global__ void __launch_bounds__(kThreads) wt_rank_kernel(Bits bits,
                                                           WtTree const * __restrict__ tree,
                                                           uint64_t size,
                                                           uint64_t sigma,
                                                           uint64_t const * __restrict__ qi,
                                                           uint8_t const * __restrict__ qc,
                                                           uint64_t n,
                                                           uint64_t * __restrict__ out)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WtTree * t = reinterpret_cast<WtTree *>(smem_raw);
    stage_tree(tree, t);
    bits.attach(smem_raw + sizeof(WtTree));
    uint64_t const stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += stride)
    {
        uint64_t i = ld_stream_u64(qi + q);
        uint32_t c = qc[q];
        uint64_t r = SDSLGPU_NPOS;
        if (i <= size)
            r = wt_rank_one(bits, t, sigma, i, c);
        st_stream_u64(out + q, r);
    }
}

template <class Bits>
__global__ void __launch_bounds__(kThreads) wt_access_kernel(Bits bits,
                                                             WtTree const * __restrict__ tree,
                                                             uint64_t size,
                                                             uint64_t const * __restrict__ qi,
                                                             uint64_t n,
                                                             uint64_t * __restrict__ sym_out,
                                                             uint64_t * __restrict__ rank_out)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WtTree * t = reinterpret_cast<WtTree *>(smem_raw);
    stage_tree(tree, t);
    bits.attach(smem_raw + sizeof(WtTree));
    uint64_t const stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += stride)
    {
        uint64_t i = ld_stream_u64(qi + q);
        uint64_t r = SDSLGPU_NPOS, s = SDSLGPU_NPOS;
        if (i < size)
        {
            uint32_t sym;
            r = wt_inverse_select_one(bits, t, i, sym);
            s = sym;
        }
        st_stream_u64(sym_out + q, s);
        if (rank_out)
            st_stream_u64(rank_out + q, r);
    }
}

// select(i, c) (wt_pc.hpp:443-474): climb from the leaf; at each parent one select on m_bv
template <class Bits>
__global__ void __launch_bounds__(kThreads) wt_select_kernel(Bits bits,
                                                             WtTree const * __restrict__ tree,
                                                             uint64_t size,
                                                             uint64_t sigma,
                                                             uint64_t const * __restrict__ qi,
                                                             uint8_t const * __restrict__ qc,
                                                             uint64_t n,
                                                             uint64_t * __restrict__ out)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WtTree * t = reinterpret_cast<WtTree *>(smem_raw);
    stage_tree(tree, t);
    bits.attach(smem_raw + sizeof(WtTree));
    uint64_t const stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += stride)
    {
        uint64_t i = ld_stream_u64(qi + q);
        uint32_t c = qc[q];
        uint32_t v = t->c_to_leaf[c];
        uint64_t r;
        if (v == kUndef)
            r = size; // the reference returns size() for an absent symbol (:447-450)
        else if (i == 0)
            r = SDSLGPU_NPOS;
        else if (sigma == 1)
            r = (i - 1 < size) ? i - 1 : size;
        else
        {
            uint64_t p = t->path[c];
            uint32_t len = (uint32_t)(p >> 56);
            uint64_t occ = t->occ[c]; // the reference leaves i > rank(size, c) undefined; here it is NPOS
            if (i > occ)
                r = SDSLGPU_NPOS;
            else
            {
                r = i - 1;
                p <<= (64 - len);
                for (uint32_t l = 0; l < len; ++l, p <<= 1)
                {
                    v = t->parent[v];
                    if ((p & 0x8000000000000000ULL) == 0)
                        r = bits.template select<0>(t->bv_pos[v] - t->bv_pos_rank[v] + r + 1) - t->bv_pos[v];
                    else
                        r = bits.template select<1>(t->bv_pos_rank[v] + r + 1) - t->bv_pos[v];
                }
            }
        }
        st_stream_u64(out + q, r);
    }
}

// ------------------------------------------------------------------------------------------------
// host entry points
// ------------------------------------------------------------------------------------------------
int wt_huff_upload(sdslgpu_handle * h, uint64_t size, uint64_t sigma, WtTree const & tree, uint64_t const * bv_words, uint64_t bv_bits, cudaStream_t s)
{
    WtHuffImage & w = h->wt;
    w.use_rrr = (h->flags & SDSLGPU_F_RRR_BV) != 0;
    if (w.use_rrr)
        SG_TRY(rrr_build_image(h->pool, w.rrr, bv_words, false, bv_bits, s));
    else
        SG_TRY(bv_build(h->pool, w.bv, h->flags & ~SDSLGPU_F_NO_SELECT, bv_words, false, bv_bits, s));
    return wt_huff_finish(h, size, sigma, tree, s);
}

// the bit vector image (plain or rrr) is in place: node ranks, symbol counts, device copy of the tree
int wt_huff_finish(sdslgpu_handle * h, uint64_t size, uint64_t sigma, WtTree const & tree, cudaStream_t s)
{
    WtHuffImage & w = h->wt;
    w.size = size;
    w.sigma = sigma;
    w.host_tree = tree;
    // inner nodes: bv_pos_rank = rank1(m_bv, bv_pos) (wt_helper.hpp:319-327), computed on the device
    uint32_t nn = tree.nnodes;
    if (nn)
    {
        std::vector<uint64_t> pos(nn), rk(nn);
        for (uint32_t v = 0; v < nn; ++v)
            pos[v] = tree.bv_pos[v];
        uint64_t *d_pos = nullptr, *d_rk = nullptr;
        SG_TRY(h->pool.alloc_t(&d_pos, nn));
        SG_TRY(h->pool.alloc_t(&d_rk, nn));
        SG_CUDA(cudaMemcpyAsync(d_pos, pos.data(), nn * 8, cudaMemcpyHostToDevice, s));
        if (w.use_rrr)
            SG_TRY(rrr_rank_image(w.rrr, 1, d_pos, nn, d_rk, s));
        else
            SG_TRY(bv_rank_device(w.bv, 0, 1, d_pos, nn, d_rk, s));
        SG_CUDA(cudaMemcpyAsync(rk.data(), d_rk, nn * 8, cudaMemcpyDeviceToHost, s));
        SG_CUDA(cudaStreamSynchronize(s));
        h->pool.release(d_pos);
        h->pool.release(d_rk);
        WtTree & ht = w.host_tree;
        for (uint32_t v = 0; v < nn; ++v)
            if (ht.child[v][0] != kUndef)
                ht.bv_pos_rank[v] = rk[v];
        // occurrences per symbol = size of the leaf's side of its parent's bit range; nodes lie in BFS =
        // bit-vector order, so node v spans [bv_pos[v], bv_pos[v+1]) and holds rk[v+1] - rk[v] ones
        for (int c = 0; c < 256; ++c)
        {
            uint16_t v = ht.c_to_leaf[c];
            ht.occ[c] = 0;
            if (v == kUndef)
                continue;
            if (v == 0)
            {
                ht.occ[c] = size;
                continue;
            }
            uint16_t par = ht.parent[v];
            uint64_t span = ht.bv_pos[par + 1] - ht.bv_pos[par], ones = rk[par + 1] - rk[par];
            ht.occ[c] = (ht.child[par][1] == v) ? ones : span - ones;
        }
    }
    SG_TRY(h->pool.alloc_t(&w.tree, 1));
    SG_CUDA(cudaMemcpyAsync(w.tree, &w.host_tree, sizeof(WtTree), cudaMemcpyHostToDevice, s));
    SG_CUDA(cudaStreamSynchronize(s));
    return SDSLGPU_OK;
}

// Device path of the constructor: text -> HBM, symbol counts, (host: Huffman shape), bit planes (wt_build.cu), rank
// blocks + select samples (bv.cu) without the bits ever visiting the host.  *done = false: not enough device memory
// for the sort scratch (6 bytes per symbol) or SDSLGPU_HOST_WT=1 — the caller takes the host fill below.
static int wt_huff_build_on_device(sdslgpu_handle * h, uint8_t const * text, uint64_t n, cudaStream_t s, bool * done)
{
    *done = false;
    char const * force_host = std::getenv("SDSLGPU_HOST_WT");
    if (n == 0 || (force_host && std::atoi(force_host) != 0))
        return SDSLGPU_OK;
    uint8_t * d_text = nullptr;
    if (cudaMalloc(reinterpret_cast<void **>(&d_text), n) != cudaSuccess)
    {
        cudaGetLastError();
        return SDSLGPU_OK;
    }
    struct Free
    {
        void * p;
        ~Free()
        {
            cudaFree(p);
        }
    } free_text{d_text};
    SG_CUDA(cudaMemcpyAsync(d_text, text, n, cudaMemcpyHostToDevice, s));
    uint64_t C[256];
    int st = wt_histogram_device(d_text, n, C, s);
    if (st == SDSLGPU_ENOTSUP)
        return SDSLGPU_OK;
    SG_TRY(st);
    WtTree tree;
    uint64_t sigma = 0;
    uint64_t bits = build_huff_tree(C, tree, sigma);
    if (bits == kWtTooDeep)
    {
        set_error("wt_huff: Huffman code deeper than 56 levels (the reference throws \"Code depth greater than 56!!!\", wt_helper.hpp:304-307)");
        return SDSLGPU_EINVAL;
    }
    uint64_t * d_words = nullptr;
    if (cudaMalloc(reinterpret_cast<void **>(&d_words), (((bits + 63) >> 6) + 2) * 8) != cudaSuccess)
    {
        cudaGetLastError();
        return SDSLGPU_OK;
    }
    Free free_words{d_words};
    st = wt_huff_planes_device(d_text, n, tree, bits, d_words, s);
    if (st == SDSLGPU_ENOTSUP)
        return SDSLGPU_OK;
    SG_TRY(st);
    WtHuffImage & w = h->wt;
    w.use_rrr = (h->flags & SDSLGPU_F_RRR_BV) != 0;
    if (w.use_rrr)
        SG_TRY(rrr_build_image(h->pool, w.rrr, d_words, true, bits, s));
    else
        SG_TRY(bv_build(h->pool, w.bv, h->flags & ~SDSLGPU_F_NO_SELECT, d_words, true, bits, s));
    SG_TRY(wt_huff_finish(h, n, sigma, tree, s));
    *done = true;
    return SDSLGPU_OK;
}

int wt_huff_build_from_text(sdslgpu_handle * h, uint8_t const * text, uint64_t n, cudaStream_t s)
{
    bool done = false;
    SG_TRY(wt_huff_build_on_device(h, text, n, s, &done));
    if (done)
        return SDSLGPU_OK;
    uint64_t C[256] = {0};
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
                for (uint64_t k = lo; k < hi; ++k)
                    ++part[t][text[k]];
            });
        for (auto & x : th)
            x.join();
        for (unsigned t = 0; t < T; ++t)
            for (int c = 0; c < 256; ++c)
                C[c] += part[t][c];
    }
    WtTree tree;
    uint64_t sigma = 0, bits = 0;
    std::vector<uint64_t> bv(1, 0);
    if (n)
    {
        bits = build_huff_tree(C, tree, sigma);
        if (bits == kWtTooDeep)
        {
            set_error("wt_huff: Huffman code deeper than 56 levels (the reference throws \"Code depth greater than 56!!!\", wt_helper.hpp:304-307)");
            return SDSLGPU_EINVAL;
        }
        bv.assign(((bits + 63) >> 6) + 1, 0);
        fill_bit_planes(text, n, tree, bv);
    }
    else
    {
        std::memset(&tree, 0, sizeof(tree));
        for (int c = 0; c < 256; ++c)
            tree.c_to_leaf[c] = kUndef;
    }
    return wt_huff_upload(h, n, sigma, tree, bv.data(), bits, s);
}

static size_t const kTreeSmem = sizeof(WtTree);

// ------------------------------------------------------------------------------------------------
// level-synchronous rank: ALL queries advance one tree level per launch.  The bits of one depth of the tree are
// a contiguous slice of m_bv (nodes are laid out in BFS order) — 1/8 of the index for a byte alphabet — so each
// pass gathers from a slice that stays resident in the 126 MB L2 (160 G gathers/s measured) instead of
// from DRAM at random (37 G/s).  Per-query state between passes (running rank: 48 bits, current node: 9 bits)
// lives in the output array itself.
// ------------------------------------------------------------------------------------------------
static constexpr uint64_t kStateMask = (1ull << 48) - 1;

template <class Bits>
__global__ void __launch_bounds__(kThreads) wt_rank_level_kernel(Bits bits,
                                                                 WtTree const * __restrict__ tree,
                                                                 uint64_t size,
                                                                 uint32_t level,
                                                                 uint32_t last_level,
                                                                 uint64_t const * __restrict__ qi,
                                                                 uint8_t const * __restrict__ qc,
                                                                 uint64_t n,
                                                                 uint64_t * __restrict__ out)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WtTree * t = reinterpret_cast<WtTree *>(smem_raw);
    stage_tree(tree, t);
    bits.attach(smem_raw + sizeof(WtTree));
    uint64_t const stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += stride)
    {
        uint32_t c = qc[q];
        uint64_t r;
        uint32_t v = 0;
        if (level == 0)
        {
            r = qi[q];
            if (r > size)
            {
                out[q] = SDSLGPU_NPOS;
                continue;
            }
            if (t->c_to_leaf[c] == kUndef)
                r = 0;
        }
        else
        {
            uint64_t st = out[q];
            if (st == SDSLGPU_NPOS)
                continue;
            r = st & kStateMask;
            v = (uint32_t)(st >> 48);
        }
        uint64_t p = t->path[c];
        uint32_t len = (uint32_t)(p >> 56);
        if (level < len && r != 0)
        {
            uint32_t bit = (uint32_t)(p >> level) & 1u;
            uint64_t o = bits.rank1(t->bv_pos[v] + r) - t->bv_pos_rank[v];
            r = bit ? o : r - o;
            v = t->child[v][bit];
        }
        out[q] = (level == last_level) ? r : (r | ((uint64_t)v << 48));
    }
}

// level-synchronous operator[] / inverse_select (wt_pc.hpp:336-357, 411-430): every query descends one depth per
// launch; state (position inside the node: 48 bits, node: 16 bits) lives in `state` (= rank_out, or sym_out when the
// caller wants symbols only).  Queries that reach their leaf early (short Huffman codes) just wait.
template <class Bits>
__global__ void __launch_bounds__(kThreads) wt_access_level_kernel(Bits bits,
                                                                   WtTree const * __restrict__ tree,
                                                                   uint64_t size,
                                                                   uint32_t level,
                                                                   uint32_t last_level,
                                                                   uint64_t const * __restrict__ qi,
                                                                   uint64_t n,
                                                                   uint64_t * __restrict__ state,
                                                                   uint64_t * __restrict__ sym_out,
                                                                   uint64_t * __restrict__ rank_out)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WtTree * t = reinterpret_cast<WtTree *>(smem_raw);
    stage_tree(tree, t);
    bits.attach(smem_raw + sizeof(WtTree));
    uint64_t const stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += stride)
    {
        uint64_t i;
        uint32_t v = 0;
        if (level == 0)
        {
            i = qi[q];
            if (i >= size)
            {
                sym_out[q] = SDSLGPU_NPOS;
                if (rank_out)
                    rank_out[q] = SDSLGPU_NPOS;
                continue;
            }
        }
        else
        {
            uint64_t st = state[q];
            if (st == SDSLGPU_NPOS)
                continue;
            i = st & kStateMask;
            v = (uint32_t)(st >> 48);
        }
        if (t->child[v][0] != kUndef)
        {
            uint32_t bit;
            uint64_t o = bits.rank1_and_bit(t->bv_pos[v] + i, bit) - t->bv_pos_rank[v];
            i = bit ? o : i - o;
            v = t->child[v][bit];
        }
        if (level == last_level)
        { // every path has ended by now
            sym_out[q] = t->bv_pos_rank[v];
            if (rank_out)
                rank_out[q] = i;
        }
        else
            state[q] = i | ((uint64_t)v << 48);
    }
}

// level-synchronous select(i, c) (wt_pc.hpp:443-474): the climb from the leaf, one ABSOLUTE tree depth per launch
// (deepest first), so that all selects of a launch fall into the bit range of one depth.  A symbol whose code has
// `len` bits takes its first step in the launch for depth len-1.
template <class Bits>
__global__ void __launch_bounds__(kThreads) wt_select_level_kernel(Bits bits,
                                                                   WtTree const * __restrict__ tree,
                                                                   uint64_t size,
                                                                   uint32_t level,
                                                                   uint32_t top_level,
                                                                   uint64_t const * __restrict__ qi,
                                                                   uint8_t const * __restrict__ qc,
                                                                   uint64_t n,
                                                                   uint64_t * __restrict__ out)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WtTree * t = reinterpret_cast<WtTree *>(smem_raw);
    stage_tree(tree, t);
    bits.attach(smem_raw + sizeof(WtTree));
    uint64_t const stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += stride)
    {
        uint32_t c = qc[q];
        uint64_t p = t->path[c];
        uint32_t len = (uint32_t)(p >> 56);
        uint32_t leaf = t->c_to_leaf[c];
        if (level == top_level)
        { // first launch: settle the answers that need no climbing and the out-of-domain ones
            uint64_t i = qi[q];
            if (leaf == kUndef)
            {
                out[q] = size; // the reference returns size() for an absent symbol (wt_pc.hpp:447-450)
                continue;
            }
            if (i == 0 || i > t->occ[c])
            {
                out[q] = SDSLGPU_NPOS;
                continue;
            }
            if (len <= level)
            { // starts in a later launch (shorter code): park the state
                out[q] = (i - 1) | ((uint64_t)leaf << 48);
                continue;
            }
        }
        uint64_t r;
        uint32_t v;
        if (level == top_level)
        { // the deepest codes take their first step in the first launch
            r = qi[q] - 1;
            v = leaf;
        }
        else
        {
            if (leaf == kUndef || len <= level)
                continue; // settled (absent symbol), or not started yet
            uint64_t st = out[q];
            if (st == SDSLGPU_NPOS)
                continue; // out of domain, settled in the first launch
            r = st & kStateMask;
            v = (uint32_t)(st >> 48);
        }
        uint32_t bit = (uint32_t)(p >> level) & 1u; // path bit taken at depth `level` (LSB first from the root)
        v = t->parent[v];
        if (bit == 0)
            r = bits.template select<0>(t->bv_pos[v] - t->bv_pos_rank[v] + r + 1) - t->bv_pos[v];
        else
            r = bits.template select<1>(t->bv_pos_rank[v] + r + 1) - t->bv_pos[v];
        out[q] = (level == 0) ? r : (r | ((uint64_t)v << 48));
    }
}

static uint32_t wt_depth(WtHuffImage const & w)
{
    uint32_t depth = 0;
    for (int k = 0; k < 256; ++k)
        if (w.host_tree.c_to_leaf[k] != kUndef)
            depth = std::max(depth, (uint32_t)(w.host_tree.path[k] >> 56));
    return depth;
}

// one launch per tree depth pays when the batch is large and a depth's share of m_bv can stay in the L2
static bool wt_level_sync(WtHuffImage const & w, uint64_t n, uint32_t depth)
{
    bool level_sync = n >= (1u << 16) && depth >= 2 && depth <= 24 && w.sigma > 1 && (w.use_rrr ? w.rrr.size : w.bv.nbits) < (1ull << 47);
    if (char const * e = std::getenv("SDSLGPU_WT_LEVEL_SYNC")) // tuning knob for experiments
        level_sync = std::atoi(e) != 0 && depth >= 1 && w.sigma > 1;
    return level_sync;
}

int wt_rank_device(sdslgpu_handle const * h, uint64_t const * i, uint8_t const * c, uint64_t n, uint64_t * out, cudaStream_t s)
{
    WtHuffImage const & w = h->wt;
    if (n == 0)
        return SDSLGPU_OK;
    uint32_t depth = wt_depth(w); // = number of passes of the level-synchronous form
    if (!wt_level_sync(w, n, depth))
    {
        SG_LAUNCH_BITS(wt_rank_kernel, w, grid_for(n), kTreeSmem, s, w.tree, w.size, w.sigma, i, c, n, out);
        SG_CUDA(cudaGetLastError());
        return SDSLGPU_OK;
    }
    for (uint32_t l = 0; l < depth; ++l)
    {
        SG_LAUNCH_BITS(wt_rank_level_kernel, w, grid_for(n), kTreeSmem, s, w.tree, w.size, l, depth - 1, i, c, n, out);
        SG_CUDA(cudaGetLastError());
    }
    return SDSLGPU_OK;
}

int wt_select_device(sdslgpu_handle const * h, uint64_t const * i, uint8_t const * c, uint64_t n, uint64_t * out, cudaStream_t s)
{
    WtHuffImage const & w = h->wt;
    if (n == 0)
        return SDSLGPU_OK;
    uint32_t depth = wt_depth(w);
    if (wt_level_sync(w, n, depth))
    {
        for (uint32_t l = depth; l-- > 0;)
        {
            SG_LAUNCH_BITS(wt_select_level_kernel, w, grid_for(n), kTreeSmem, s, w.tree, w.size, l, depth - 1, i, c, n, out);
            SG_CUDA(cudaGetLastError());
        }
        return SDSLGPU_OK;
    }
    SG_LAUNCH_BITS(wt_select_kernel, w, grid_for(n), kTreeSmem, s, w.tree, w.size, w.sigma, i, c, n, out);
    SG_CUDA(cudaGetLastError());
    return SDSLGPU_OK;
}

int wt_access_device(sdslgpu_handle const * h, uint64_t const * i, uint64_t n, uint64_t * sym, uint64_t * rnk, cudaStream_t s)
{
    WtHuffImage const & w = h->wt;
    if (n == 0)
        return SDSLGPU_OK;
    uint32_t depth = wt_depth(w);
    if (wt_level_sync(w, n, depth))
    {
        uint64_t * state = rnk ? rnk : sym;
        for (uint32_t l = 0; l < depth; ++l)
        {
            SG_LAUNCH_BITS(wt_access_level_kernel, w, grid_for(n), kTreeSmem, s, w.tree, w.size, l, depth - 1, i, n, state, sym, rnk);
            SG_CUDA(cudaGetLastError());
        }
        return SDSLGPU_OK;
    }
    SG_LAUNCH_BITS(wt_access_kernel, w, grid_for(n), kTreeSmem, s, w.tree, w.size, i, n, sym, rnk);
    SG_CUDA(cudaGetLastError());
    return SDSLGPU_OK;
}

} // namespace sdslgpu
