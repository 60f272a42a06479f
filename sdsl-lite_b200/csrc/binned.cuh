// binned.cuh — locality-ordered execution of very large query batches whose answer lives near `key` in the index
// (rank: key = position; select: key = i - 1, the i-th one sits at a position that grows with i).
//
// Same results as the one-thread-per-query kernels, different order of work.  A B200 serves ~38-44 G random DRAM
// gathers/s whatever their width but several times that from the 126 MB L2.  For batches of tens of
// millions of queries over an index larger than L2 the batch is therefore re-ordered so that all queries in flight
// touch one L2-sized chunk of the index:
//
//   1. bin_tile_sort_kernel   one CTA per tile of 8192 queries: counting sort of the tile by bin = key >> shift
//                             (a bin = a contiguous 16 - 24 MB chunk of the index, per op: bin_chunk_mib).  Writes, per tile, the 32-bit in-bin
//                             offsets in bin order (`recs`), every query's slot (`lp`, u16) and the tile's bin
//                             boundaries (`loff`, u16).  Streaming: 8 B read + 6 B written per query.
//   2. bin_apply_kernel<Op>   warps take (bin, tile) runs from a global ticket counter in BIN-MAJOR order, so at any
//                             moment the whole GPU gathers from one or two chunks: L2 hits after the first touch, and
//                             the index is read from DRAM once per batch.  Results go to the runs' own slots.
//   3. bin_unsort_kernel      one CTA per tile: the tile's results (contiguous) -> shared memory -> caller's order
//                             (multi-GPU group calls: and, with the same coalesced stores, into the result arrays of
//                             every other member of the group over NVLink — the all-gather fused into this stage).
//
// No global sort, no scan across tiles; deterministic results.  ~22 B per query of extra coalesced streams next to
// the 16 B of query + result every path moves.  (Storing rank results as u32 differences to the bin's first rank
// halves two of those streams but the un-sort then needs the bin of every slot; measured, it gains nothing:
// profiles/r01d_binned_v4_relative_*; so do u32 answers with bit planes for the high bits, profiles/r02s_*.)
// The speed of the scheme comes from the number of queries per index line IN ONE BATCH: a batch cut into pieces —
// sub-batches on two streams to hide a fused NVLink gather, profiles/r02z_fan_sub_batches_n2.jsonl — loses it.
#pragma once
#include "internal.h"

namespace sdslgpu
{

static constexpr int kTile = 8192;        // queries per tile (slots fit u16)
static constexpr int kTileThreads = 1024; // CTA size of the sort / un-sort kernels
static constexpr uint32_t kMaxBins = 254; // valid bins; bin index nb collects the out-of-domain queries
static constexpr uint32_t kTicketLanes = 32;  // independent ticket counters (128 bytes apart), run w belongs to counter w % 32
static constexpr uint32_t kTicketStride = 16; // in u64 units

struct BinPlan
{
    uint32_t shift = 0, nb = 0;
    uint64_t ntiles = 0;
};

// scratch of one call, carved from a single stream-ordered allocation
struct BinScratch
{
    uint8_t * mem = nullptr;
    cudaStream_t s = nullptr;
    uint32_t * recs = nullptr;
    uint16_t * lp = nullptr;
    uint16_t * loff = nullptr;
    uint64_t * res = nullptr;
    unsigned long long * ticket = nullptr;
    ~BinScratch()
    {
        if (mem)
            cudaFreeAsync(mem, s);
    }
};

// binned.cu
bool bin_make_plan(uint64_t index_bytes, uint64_t maxkey, uint64_t n, BinPlan & p, uint32_t chunk_mib = 24);
int bin_scratch_alloc(BinScratch & w, BinPlan const & p, cudaStream_t s);
int bin_launch_tile_sort(BinPlan const & p, BinScratch const & w, uint64_t const * q, uint64_t n, uint64_t sub, uint64_t maxkey, bool clamp, cudaStream_t s);
int bin_launch_unsort(BinPlan const & p, BinScratch const & w, uint64_t n, uint64_t * out, cudaStream_t s, Fan const * fan = nullptr);
unsigned bin_apply_grid(BinPlan const & p);

// Op: plain-old-data functor with
//   static constexpr int kIlp          independent gathers per lane and trip
//   static constexpr int kMinCtas      resident CTAs per SM the kernel is compiled for (register budget)
//   static constexpr uint32_t kSmem    bytes of dynamic shared memory its tables need (0: none)
//   __device__ void stage(uint8_t *)   copy tables into shared memory (called by every thread; must __syncthreads if kSmem)
//   __device__ uint64_t operator()(uint64_t key) const
//   static constexpr int kLookAhead    (optional; default 2) how far bin_apply_kernel issues memory operations ahead of
//                                      their use: 0 nothing, 1 the next run's ticket, 2 also the next trip's record —
//                                      each costs registers, so ops at the edge of their register budget opt out
//   static constexpr uint32_t kChunkMiB (optional; default 24) the share of the index one bin covers
template <class Op, class = void>
struct bin_chunk_mib
{
    static constexpr uint32_t value = 24;
};
template <class Op>
struct bin_chunk_mib<Op, decltype((void)Op::kChunkMiB)>
{
    static constexpr uint32_t value = Op::kChunkMiB;
};
template <class Op, class = void>
struct bin_look_ahead
{
    static constexpr int value = 2;
};
template <class Op>
struct bin_look_ahead<Op, decltype((void)Op::kLookAhead)>
{
    static constexpr int value = Op::kLookAhead;
};

template <class Op>
__global__ void __launch_bounds__(kThreads, Op::kMinCtas) bin_apply_kernel(Op op,
                                                             uint32_t const * __restrict__ recs,
                                                             uint16_t const * __restrict__ loff,
                                                             uint32_t shift,
                                                             uint32_t nb,
                                                             uint64_t ntiles,
                                                             unsigned long long * __restrict__ ticket,
                                                             uint64_t * __restrict__ res)
{
    extern __shared__ __align__(16) uint8_t bin_smem[];
    op.stage(bin_smem);
    uint32_t const lane = threadIdx.x & 31u;
    uint64_t const runs = (uint64_t)nb * ntiles;
    // Runs are handed out by ticket counters in increasing order, so the runs in flight are always one contiguous
    // window of the bin-major sequence (about one run per resident warp: less than one bin) whatever the residency
    // or the speed of individual warps.  32 counters, each owning the runs w = 32 k + c, keep the atomics off one address.
    uint32_t const c = (blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5)) % kTicketLanes;
    // A warp's work is a chain of dependent memory operations: ticket -> bin boundaries -> query record -> the op's own
    // gathers.  Two of the links are taken off the critical path by issuing them one step ahead: the NEXT run's ticket
    // is drawn while the current run is answered, and the NEXT trip's record is loaded before the current trip's gathers.
    auto draw = [&]() -> unsigned long long {
        unsigned long long w = 0;
        if (lane == 0)
            w = atomicAdd(ticket + c * kTicketStride, 1ull);
        return w;
    };
    constexpr int kAhead = bin_look_ahead<Op>::value;
    unsigned long long drawn = kAhead >= 1 ? draw() : 0ull;
    for (;;)
    {
        if (kAhead < 1)
            drawn = draw();
        unsigned long long const w0 = __shfl_sync(0xFFFFFFFFu, drawn, 0) * kTicketLanes + c;
        if (w0 >= runs)
            break;
        if (kAhead >= 1)
            drawn = draw(); // in flight until the next round of this loop
        {
            uint64_t const w = w0;
            uint32_t const b = (uint32_t)(w / ntiles);
            uint64_t const t = w - (uint64_t)b * ntiles;
            uint16_t const * o = loff + t * (nb + 2) + b;
            uint32_t const o0 = __ldg(o), o1 = __ldg(o + 1);
            uint64_t const hi = (uint64_t)b << shift;
            uint32_t const * r_in = recs + t * kTile;
            uint64_t * r_out = res + t * kTile;
            constexpr int I = Op::kIlp;
            if (I == 1 && kAhead >= 2)
            {
                uint32_t k = o0 + lane;
                uint32_t rec = k < o1 ? ld_stream_u32(r_in + k) : 0u;
                while (k < o1)
                {
                    uint32_t const kn = k + 32;
                    uint32_t const recn = kn < o1 ? ld_stream_u32(r_in + kn) : 0u; // next trip's record: no consumer yet
                    st_stream_u64(r_out + k, op(hi + rec));
                    k = kn;
                    rec = recn;
                }
                continue;
            }
            for (uint32_t k = o0 + lane; k < o1; k += 32 * I)
            { // I independent gathers per lane and trip
                uint64_t key[I], a[I];
#pragma unroll
                for (int u = 0; u < I; ++u)
                    key[u] = hi + ld_stream_u32(r_in + ((k + 32 * u < o1) ? k + 32 * u : k));
#pragma unroll
                for (int u = 0; u < I; ++u)
                    a[u] = op(key[u]);
#pragma unroll
                for (int u = 0; u < I; ++u)
                    if (k + 32 * u < o1)
                        st_stream_u64(r_out + k + 32 * u, a[u]);
            }
        }
    }
}

// the whole pipeline.  key = q - sub, in domain iff key <= maxkey; out-of-domain queries get SDSLGPU_NPOS
// (clamp_high: keys past maxkey are answered as maxkey, only a wrapped-around 0 - sub is out of domain).
// *done = false (and nothing launched) when the plan does not fit (shift > 32, tile count overflow).
template <class Op>
int bin_run(Op const & op, uint64_t index_bytes, uint64_t sub, uint64_t maxkey, uint64_t const * q, uint64_t n, uint64_t * out, cudaStream_t s, bool * done,
            bool clamp_high = false, Fan const * fan = nullptr)
{
    *done = false;
    BinPlan p;
    if (n == 0 || !bin_make_plan(index_bytes, maxkey, n, p, bin_chunk_mib<Op>::value))
        return SDSLGPU_OK;
    BinScratch w;
    SG_TRY(bin_scratch_alloc(w, p, s));
    SG_TRY(bin_launch_tile_sort(p, w, q, n, sub, maxkey, clamp_high, s));
    if (Op::kSmem > 48 * 1024)
        SG_CUDA(cudaFuncSetAttribute(bin_apply_kernel<Op>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Op::kSmem));
    bin_apply_kernel<Op><<<bin_apply_grid(p), kThreads, Op::kSmem, s>>>(op, w.recs, w.loff, p.shift, p.nb, p.ntiles, w.ticket, w.res);
    SG_CUDA(cudaGetLastError());
    SG_TRY(bin_launch_unsort(p, w, n, out, s, fan));
    *done = true;
    return SDSLGPU_OK;
}

} // namespace sdslgpu
