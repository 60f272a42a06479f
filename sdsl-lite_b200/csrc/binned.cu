// binned.cu — the op-independent parts of the locality-ordered batch pipeline (binned.cuh): tile counting sort,
// un-sort, plan / scratch, and the plain bit vector's rank / select ops on top of it
// (rank_support_v.hpp:129-139, select_support_mcl.hpp:384-439 — same results, different order of work).
#include <cstdlib>

#include "binned.cuh"
#include "bv_device.cuh"
#include "fan.cuh"

namespace sdslgpu
{

static constexpr int kPer = kTile / kTileThreads;

// ------------------------------------------------------------------------------------------------
// 1. tile-local counting sort by bin
//    key = q - sub (sub = 0 for rank, 1 for select); valid iff key <= maxkey.
//    Persistent CTAs (two per SM); the 64 KB of keys of a CTA's NEXT tile are fetched by one TMA bulk copy
//    (cp.async.bulk + mbarrier) while the current tile is scanned, scattered and written out, so DRAM reads never
//    pause for the barrier phases.  kTma = false: plain loads (key array not 16-byte aligned).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(void const * p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void tma_load_1d(uint32_t dst, void const * src, uint32_t bytes, uint32_t bar)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    do
    {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok)
                     : "r"(bar), "r"(parity)
                     : "memory");
    } while (!ok);
}

// key j of the tile: from the TMA buffer (complete tiles) or straight from global memory
// (clamp: keys beyond maxkey — but not a wrapped-around 0 - sub — count as maxkey: for ops whose answer is the same
//  for every query past the end, like select_support_rrr's in-band size())
template <bool kFull, bool kClamp>
__device__ __forceinline__ bool tile_key(uint64_t const * __restrict__ src, uint32_t limit, uint32_t j, uint64_t sub, uint64_t maxkey, uint64_t & key)
{
    if (!kFull && j >= limit)
        return false;
    key = (kFull ? src[j] : ld_stream_u64(src + j)) - sub;
    if (kClamp && key > maxkey && key != ~0ull)
        key = maxkey;
    return true;
}

// pass A: histogram of the tile over the bins (bin nb = out of domain).  The pass keeps what pass B needs of every key
// (its 32-bit in-bin offset and its bin, a byte) in registers — 10 registers across two barriers — so the key buffer
// is free, and the NEXT tile's TMA copy under way, as soon as this pass is over, and pass B reads no key a second time
// (round 2: sort stage 0.296 -> 0.27 ms per 1e8 keys, profiles/r02x_*; re-reading the keys from shared memory in pass B
// had been round 1's way of living with 32 registers while the keys themselves were kept).
template <bool kFull, bool kClamp>
__device__ __forceinline__ void tile_count_keep(uint64_t const * __restrict__ src, uint32_t limit, uint64_t sub, uint64_t maxkey, uint32_t shift, uint32_t mask32,
                                                uint32_t nb, uint32_t * __restrict__ cnt, uint32_t tid, uint32_t (&rec)[kPer], uint32_t (&bins)[(kPer + 3) / 4])
{
#pragma unroll
    for (int u = 0; u < (kPer + 3) / 4; ++u)
        bins[u] = 0;
#pragma unroll
    for (int u = 0; u < kPer; ++u)
    {
        uint64_t key = 0;
        bool const have = tile_key<kFull, kClamp>(src, limit, (uint32_t)u * kTileThreads + tid, sub, maxkey, key);
        uint32_t const b = (key <= maxkey) ? (uint32_t)(key >> shift) : nb; // nb <= kMaxBins = 254: a byte
        rec[u] = (uint32_t)key & mask32;
        bins[u >> 2] |= b << (8 * (u & 3));
        if (have)
            atomicAdd(&cnt[b], 1u);
    }
}
// pass B: every key takes the next free slot of its bin (cur[] starts at the bins' exclusive offsets): in-bin offsets
// to shared memory in slot order, slots to `lp`
template <bool kFull>
__device__ __forceinline__ void tile_scatter_kept(uint32_t limit, uint32_t const (&rec)[kPer], uint32_t const (&bins)[(kPer + 3) / 4], uint32_t * __restrict__ cur,
                                                  uint32_t * __restrict__ srec, uint16_t * __restrict__ lp_t, uint32_t tid)
{
#pragma unroll
    for (int u = 0; u < kPer; ++u)
    {
        if (kFull || (uint32_t)u * kTileThreads + tid < limit)
        {
            uint32_t l = atomicAdd(&cur[(bins[u >> 2] >> (8 * (u & 3))) & 0xFFu], 1u);
            srec[l] = rec[u];
            lp_t[u * kTileThreads] = (uint16_t)l;
        }
    }
}

template <bool kTma, bool kClamp>
__global__ void __launch_bounds__(kTileThreads, 2) bin_tile_sort_kernel(uint64_t const * __restrict__ q,
                                                                        uint64_t n,
                                                                        uint64_t sub,
                                                                        uint64_t maxkey,
                                                                        uint32_t shift,
                                                                        uint32_t nb,
                                                                        uint64_t ntiles,
                                                                        uint32_t * __restrict__ recs,
                                                                        uint16_t * __restrict__ lp,
                                                                        uint16_t * __restrict__ loff)
{
    extern __shared__ __align__(128) uint8_t sort_smem[];
    uint32_t * srec = reinterpret_cast<uint32_t *>(sort_smem);             // kTile * 4 bytes
    uint64_t * skey = reinterpret_cast<uint64_t *>(sort_smem + kTile * 4); // kTile * 8 bytes (kTma only)
    // cnt: counts, then exclusive offsets (entry nb+1 = queries in the tile); two copies used alternately, so that
    // zeroing the next tile's counters never races with a slow thread still reading this tile's.  cur: slot cursors.
    __shared__ uint32_t cnt2[2][kMaxBins + 2];
    __shared__ uint32_t cur[kMaxBins + 2];
    __shared__ __align__(8) uint64_t bar_mem;
    uint32_t const tid = threadIdx.x;
    uint32_t const bar = smem_u32(&bar_mem), skey_addr = smem_u32(skey);
    uint32_t const mask32 = shift >= 32 ? 0xFFFFFFFFu : (1u << shift) - 1u; // shift <= 32
    uint32_t parity = 0;
    if (kTma)
    {
        if (tid == 0)
        {
            mbar_init(bar, 1);
            if ((uint64_t)(blockIdx.x + 1) * kTile <= n) // a full first tile: fetch it
                tma_load_1d(skey_addr, q + (uint64_t)blockIdx.x * kTile, kTile * 8, bar);
        }
    }
    uint32_t flip = 0;
    // The first tile's counters are zeroed here; every later tile's while the tile before it is scattered (the other
    // copy of cnt2 is idle then), so a warp that has written out its share of tile t walks straight into the count
    // pass of tile t+1: three barriers per tile instead of four.
    for (uint32_t k = tid; k < nb + 2; k += kTileThreads)
        cnt2[0][k] = 0;
    __syncthreads(); // (and the mbarrier is initialised)
    for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, flip ^= 1u)
    {
        uint32_t * const cnt = cnt2[flip];
        uint64_t const base = tile * kTile;
        bool const full = kTma && base + kTile <= n; // partial last tile: plain loads
        uint32_t const limit = (n - base < (uint64_t)kTile) ? (uint32_t)(n - base) : (uint32_t)kTile;
        uint32_t rec[kPer], bins[(kPer + 3) / 4];
        if (full)
        {
            mbar_wait(bar, parity);
            parity ^= 1u;
            tile_count_keep<true, kClamp>(skey, limit, sub, maxkey, shift, mask32, nb, cnt, tid, rec, bins);
        }
        else
            tile_count_keep<false, kClamp>(q + base, limit, sub, maxkey, shift, mask32, nb, cnt, tid, rec, bins);
        __syncthreads(); // counts complete; every key has been read out of skey
        if (kTma && tid == 0)
        {
            uint64_t next = tile + gridDim.x;
            if (next < ntiles && (next + 1) * kTile <= n)
                tma_load_1d(skey_addr, q + next * kTile, kTile * 8, bar);
        }
        if (tid < 32)
        { // exclusive scan of the nb+1 counters; entry nb+1 receives the total
            uint32_t carry = 0;
            for (uint32_t c0 = 0; c0 < nb + 2; c0 += 32)
            {
                uint32_t k = c0 + tid;
                uint32_t v = (k <= nb) ? cnt[k] : 0u, x = v;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1)
                {
                    uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, d);
                    if ((int)tid >= d)
                        x += y;
                }
                if (k < nb + 2)
                {
                    cnt[k] = carry + x - v;
                    cur[k] = carry + x - v;
                }
                carry += __shfl_sync(0xFFFFFFFFu, x, 31);
            }
        }
        __syncthreads(); // (early zero: every thread has left the write-out of the previous tile — srec is free again)
        for (uint32_t k = tid; k < nb + 2; k += kTileThreads)
        {
            loff[tile * (nb + 2) + k] = (uint16_t)cnt[k];
            cnt2[flip ^ 1u][k] = 0; // the next tile's counters: last read in the previous tile's write-out, two barriers ago
        }
        if (limit == kTile)
            tile_scatter_kept<true>(limit, rec, bins, cur, srec, lp + base + tid, tid);
        else
            tile_scatter_kept<false>(limit, rec, bins, cur, srec, lp + base + tid, tid);
        __syncthreads(); // srec complete
        uint32_t const total = cnt[nb]; // valid queries only: the out-of-domain slots are never read
        uint32_t * const recs_t = recs + base + tid;
        if (total == kTile)
        {
#pragma unroll
            for (int u = 0; u < kPer; ++u)
                recs_t[u * kTileThreads] = srec[u * kTileThreads + tid];
        }
        else
        {
            for (uint32_t k = tid; k < total; k += kTileThreads)
                recs_t[k - tid] = srec[k];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// 3. back to the caller's order
// ------------------------------------------------------------------------------------------------
// kFan = 1: every result is also stored to the same index of the other group members' result arrays (peer memory
// over NVLink): each warp's store is 256 contiguous bytes per destination.  kFan = 2: the same results as w-bit fields
// into the peers' staging regions (fan.cuh): w/2 words per warp and destination instead of 32.
template <int kFan>
__global__ void __launch_bounds__(kTileThreads, 2) bin_unsort_kernel(uint64_t const * __restrict__ res,
                                                                     uint16_t const * __restrict__ lp,
                                                                     uint16_t const * __restrict__ loff,
                                                                     uint32_t nb,
                                                                     uint64_t n,
                                                                     uint64_t * __restrict__ out,
                                                                     Fan const fan)
{
    extern __shared__ __align__(16) uint8_t unsort_smem[];
    uint64_t * sres = reinterpret_cast<uint64_t *>(unsort_smem);
    uint32_t const tid = threadIdx.x;
    uint64_t const tile = blockIdx.x, first = tile * kTile;
    uint32_t const nvalid = loff[tile * (nb + 2) + nb]; // slots >= nvalid hold the out-of-domain queries
    uint32_t l[kPer];
#pragma unroll
    for (int u = 0; u < kPer; ++u)
    {
        uint64_t p = first + (uint64_t)u * kTileThreads + tid;
        l[u] = (p < n) ? ld_stream_u16(lp + p) : 0xFFFFu;
    }
    for (uint32_t k = tid; k < nvalid; k += kTileThreads)
        sres[k] = ld_stream_u64(res + first + k);
    __syncthreads();
#pragma unroll
    for (int u = 0; u < kPer; ++u)
    {
        uint64_t p = first + (uint64_t)u * kTileThreads + tid;
        uint64_t const v = l[u] < nvalid ? sres[l[u]] : SDSLGPU_NPOS;
        if (p < n)
        {
            st_stream_u64(out + p, v);
            if (kFan == 1)
                fan_store(fan, p, v);
        }
        if (kFan == 2)
        { // warp-collective: 32 consecutive p per warp and u
            uint64_t const p0 = p - (tid & 31u);
            if (p0 < n)
                fan_store_packed(fan, p0, tid & 31u, n - p0 < 32 ? (uint32_t)(n - p0) : 32u, v);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static uint64_t chunk_target_bytes(uint32_t chunk_mib)
{
    if (char const * e = std::getenv("SDSLGPU_BIN_CHUNK_BYTES")) // test / tuning knob
    {
        long long v = std::atoll(e);
        if (v > 0)
            return (uint64_t)v;
    }
    return (uint64_t)chunk_mib << 20;
}

// bins are equal ranges of the key space sized so that one bin's share of the index is ~chunk_mib MiB (24 unless the op
// says otherwise: what is in flight at any moment — one bin's lines and the streams passing by — has to stay in the L2;
// measured on the 2^33-bit vector, profiles/r02s_*, r02t_*: rank 16 MiB 1.177 ms, 24: 1.196, 32: 1.198, 48: 1.60)
bool bin_make_plan(uint64_t index_bytes, uint64_t maxkey, uint64_t n, BinPlan & p, uint32_t chunk_mib)
{
    uint64_t target = chunk_target_bytes(chunk_mib);
    uint64_t want = (index_bytes + target - 1) / target;
    if (want < 1)
        want = 1;
    if (want > kMaxBins)
        want = kMaxBins;
    uint32_t s = 0;
    while (s < 63 && (maxkey >> s) + 1 > want)
        ++s;
    if (s > 32)
        return false;
    p.shift = s;
    p.nb = (uint32_t)((maxkey >> s) + 1);
    p.ntiles = (n + kTile - 1) / kTile;
    return p.ntiles < (1ull << 31);
}

bool bin_wanted(int order, uint64_t index_bytes, uint64_t n, uint32_t index_bytes_per_query)
{
    if (order == SDSLGPU_ORDER_BINNED)
        return true;
    if (order == SDSLGPU_ORDER_DIRECT)
        return false;
    // auto: only when the index cannot live in L2 and the batch is dense enough that the queries of a bin share cache
    // lines; a sparser batch misses in DRAM anyway and the two extra passes are pure cost.  The break-even density is
    // measured (tools/sweep_order.py, profiles/r02z_sweep_order.jsonl, 1.23 GB index): a one-gather op (rank, select
    // through select sectors) pays from 1 query per 128 bytes of index (9.6e6 queries: 0.98x at 8.4e6, 1.08 - 1.15x at
    // 1.25e7), the sampled select (sample + blocks) from 1 per 192 bytes (6.4e6: 0.85x at 4.2e6, 1.06x at 8.4e6).
    return index_bytes >= (192ull << 20) && n >= (1ull << 21) && n >= index_bytes / index_bytes_per_query;
}

static constexpr uint64_t kTicketBytes = kTicketLanes * kTicketStride * 8;

static uint64_t up256(uint64_t x)
{
    return (x + 255) & ~255ull;
}

int bin_scratch_alloc(BinScratch & w, BinPlan const & p, cudaStream_t s)
{
    uint64_t slots = p.ntiles * kTile;
    uint64_t b_recs = up256(slots * 4), b_lp = up256(slots * 2), b_loff = up256(p.ntiles * (p.nb + 2) * 2), b_res = up256(slots * 8);
    w.s = s;
    cudaError_t e = cudaMallocAsync(reinterpret_cast<void **>(&w.mem), kTicketBytes + b_recs + b_lp + b_loff + b_res, s);
    if (e != cudaSuccess)
    {
        w.mem = nullptr;
        return cuda_fail(e, "cudaMallocAsync (binned batch scratch)", __FILE__, __LINE__);
    }
    uint8_t * m = w.mem;
    w.ticket = reinterpret_cast<unsigned long long *>(m);
    m += kTicketBytes;
    w.res = reinterpret_cast<uint64_t *>(m);
    m += b_res;
    w.recs = reinterpret_cast<uint32_t *>(m);
    m += b_recs;
    w.lp = reinterpret_cast<uint16_t *>(m);
    m += b_lp;
    w.loff = reinterpret_cast<uint16_t *>(m);
    SG_CUDA(cudaMemsetAsync(w.ticket, 0, kTicketBytes, s));
    return SDSLGPU_OK;
}

template <bool kClamp>
static int launch_tile_sort(BinPlan const & p, BinScratch const & w, uint64_t const * q, uint64_t n, uint64_t sub, uint64_t maxkey, cudaStream_t s)
{
    // persistent CTAs, two per SM (96 KB of shared memory each with the TMA key buffer)
    uint64_t const resident = 2ull * (uint64_t)sm_count();
    unsigned grid = (unsigned)(p.ntiles < resident ? p.ntiles : resident);
    if ((reinterpret_cast<uintptr_t>(q) & 15u) == 0)
    {
        int smem = kTile * 4 + kTile * 8;
        SG_CUDA(cudaFuncSetAttribute(bin_tile_sort_kernel<true, kClamp>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        bin_tile_sort_kernel<true, kClamp><<<grid, kTileThreads, smem, s>>>(q, n, sub, maxkey, p.shift, p.nb, p.ntiles, w.recs, w.lp, w.loff);
    }
    else
        bin_tile_sort_kernel<false, kClamp><<<grid, kTileThreads, kTile * 4, s>>>(q, n, sub, maxkey, p.shift, p.nb, p.ntiles, w.recs, w.lp, w.loff);
    SG_CUDA(cudaGetLastError());
    return SDSLGPU_OK;
}

// clamp (select_support_rrr's in-band size() past the last bit) is a template parameter: the plain ops' sort carries no
// trace of it
int bin_launch_tile_sort(BinPlan const & p, BinScratch const & w, uint64_t const * q, uint64_t n, uint64_t sub, uint64_t maxkey, bool clamp, cudaStream_t s)
{
    return clamp ? launch_tile_sort<true>(p, w, q, n, sub, maxkey, s) : launch_tile_sort<false>(p, w, q, n, sub, maxkey, s);
}

template <int kFan>
static int launch_unsort(BinPlan const & p, BinScratch const & w, uint64_t n, uint64_t * out, cudaStream_t s, Fan const & fan)
{
    // 64 KB of dynamic shared memory for the result tile (attribute is per device / context: set on every call)
    SG_CUDA(cudaFuncSetAttribute(bin_unsort_kernel<kFan>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kTile * 8)));
    bin_unsort_kernel<kFan><<<(unsigned)p.ntiles, kTileThreads, kTile * 8, s>>>(w.res, w.lp, w.loff, p.nb, n, out, fan);
    SG_CUDA(cudaGetLastError());
    return SDSLGPU_OK;
}

int bin_launch_unsort(BinPlan const & p, BinScratch const & w, uint64_t n, uint64_t * out, cudaStream_t s, Fan const * fan)
{
    if (fan && fan->n)
        return fan->width ? launch_unsort<2>(p, w, n, out, s, *fan) : launch_unsort<1>(p, w, n, out, s, *fan);
    return launch_unsort<0>(p, w, n, out, s, Fan{});
}

unsigned bin_apply_grid(BinPlan const & p)
{
    uint64_t runs = (uint64_t)p.nb * p.ntiles;
    uint64_t want = (runs + kThreads / 32 - 1) / (kThreads / 32), cap = (uint64_t)sm_count() * 8;
    return (unsigned)(want < cap ? (want ? want : 1) : cap);
}

// ------------------------------------------------------------------------------------------------
// plain bit vector ops
// ------------------------------------------------------------------------------------------------
template <int B>
struct BvRankOp
{
    // measured (profiles/r01d_binned_variants.txt): 1 gather per lane x 8 CTAs/SM beats 2 x 6, 3 x 5 and 4 x 4 —
    // the L2 gather rate, not latency, is the limit, and more resident warps keep it fed
#ifndef BIN_RANK_ILP
#define BIN_RANK_ILP 1
#define BIN_RANK_CTAS 8
#endif
    static constexpr int kIlp = BIN_RANK_ILP;
    static constexpr int kMinCtas = BIN_RANK_CTAS;
#ifndef BIN_RANK_LOOKAHEAD
#define BIN_RANK_LOOKAHEAD 2
#endif
    static constexpr int kLookAhead = BIN_RANK_LOOKAHEAD;
    static constexpr uint32_t kChunkMiB = 16; // a one-gather op turns its bins over fastest: smaller bins, -1.6 %
    static constexpr uint32_t kSmem = 0;
    BvView v;
    __device__ __forceinline__ void stage(uint8_t *) const
    {}
    __device__ __forceinline__ uint64_t operator()(uint64_t pos) const
    {
        uint64_t blk = pos / kBlockBits;
        uint32_t rem = (uint32_t)(pos - blk * kBlockBits);
        uint32_t cnt, d[7];
        ld_block(v.blocks + blk, cnt, d); // whole-line fill: the neighbours are wanted by other queries of the bin
        uint64_t r = __ldg(v.top + (blk >> kSuperShift)) + cnt + block_prefix_popc(d, rem);
        return B ? r : pos - r;
    }
};

template <int B>
struct BvSelectOp
{
#ifndef BIN_SEL_ILP
#define BIN_SEL_ILP 1
#define BIN_SEL_CTAS 8
#endif
    static constexpr int kIlp = BIN_SEL_ILP;
    static constexpr int kMinCtas = BIN_SEL_CTAS;
#ifndef BIN_SEL_LOOKAHEAD
#define BIN_SEL_LOOKAHEAD 1 // the record look-ahead spills (32 registers at 8 CTAs / SM)
#endif
    static constexpr int kLookAhead = BIN_SEL_LOOKAHEAD;
    static constexpr uint32_t kSmem = 0;
    BvView v;
    __device__ __forceinline__ void stage(uint8_t *) const
    {}
    __device__ __forceinline__ uint64_t operator()(uint64_t key) const
    {
        return bv_select<B>(v, key + 1);
    }
};

// select through the select sectors (bv_device.cuh): one gather per query, like rank; the few queries whose sector is
// marked "does not fit" take the sampled select
template <int B>
struct BvSelectSectOp
{
    static constexpr int kIlp = 1;
    static constexpr int kMinCtas = 8;
#ifndef BIN_SECT_LOOKAHEAD
#define BIN_SECT_LOOKAHEAD 2
#endif
    static constexpr int kLookAhead = BIN_SECT_LOOKAHEAD;
    static constexpr uint32_t kSmem = 0;
    BvView v;
    __device__ __forceinline__ void stage(uint8_t *) const
    {}
    __device__ __forceinline__ uint64_t operator()(uint64_t key) const
    {
        return bv_select_any<B>(v, key + 1);
    }
};

bool bv_binned_wanted(BvImage const & v, uint64_t n, bool select, int b)
{
    bool const one_gather = !select || ((b == 0 || b == 1) && v.sect[b] != nullptr);
    return bin_wanted(v.order, v.nblocks * sizeof(bvblock), n, one_gather ? kBinRankDensity : kBinSelectDensity);
}

int bv_rank_binned_device(BvImage const & v, int b, uint64_t const * idx, uint64_t n, uint64_t * out, cudaStream_t s, bool * done, Fan const * fan)
{
    uint64_t bytes = v.nblocks * sizeof(bvblock);
    if (b)
        return bin_run(BvRankOp<1>{bv_view(v)}, bytes, 0, v.nbits, idx, n, out, s, done, false, fan);
    return bin_run(BvRankOp<0>{bv_view(v)}, bytes, 0, v.nbits, idx, n, out, s, done, false, fan);
}

int bv_select_binned_device(BvImage const & v, int b, uint64_t const * idx, uint64_t n, uint64_t * out, cudaStream_t s, bool * done, Fan const * fan)
{
    *done = false;
    uint64_t args = b ? v.ones : v.nbits - v.ones, bytes = v.nblocks * sizeof(bvblock);
    if (args == 0)
        return SDSLGPU_OK; // every query is out of domain: the direct kernel answers NPOS
    if (v.sect[b])
    { // the bins are ranges of the sector array
        uint64_t const sect_bytes = v.nsect[b] * sizeof(bvblock);
        if (b)
            return bin_run(BvSelectSectOp<1>{bv_view(v)}, sect_bytes, 1, args - 1, idx, n, out, s, done, false, fan);
        return bin_run(BvSelectSectOp<0>{bv_view(v)}, sect_bytes, 1, args - 1, idx, n, out, s, done, false, fan);
    }
    if (b)
        return bin_run(BvSelectOp<1>{bv_view(v)}, bytes, 1, args - 1, idx, n, out, s, done, false, fan);
    return bin_run(BvSelectOp<0>{bv_view(v)}, bytes, 1, args - 1, idx, n, out, s, done, false, fan);
}

} // namespace sdslgpu
