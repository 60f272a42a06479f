// group.cu — the multi-GPU data plane of the C ABI (include/sdslgpu.h, "groups"): index replicated on every GPU of
// one box, a query batch sharded over the members, the results all-gathered so that every member ends up holding
// all n answers (BASELINE.json north_star: "replicated index, NCCL all-gather of results only").
// The reference has no counterpart — its queries are scalar calls on one host thread (rank_support_v.hpp:129-139).
//
// Two ways of bringing the results together:
//   SDSLGPU_GATHER_NCCL   the shard kernels, then ncclAllGather in place on the result array.
//   SDSLGPU_GATHER_FUSED  no collective call at all: the last kernel of a member's shard stores every result into
//                         its own result array AND into the result arrays of all other members (peer memory mapped
//                         by cudaIpc* between processes, by peer access inside one process), so the transfer over
//                         NVLink / NVSwitch happens tile by tile while the remaining tiles are still being computed.
//                         For plain bit vectors that kernel is the un-sort stage of the binned pipeline
//                         (bin_unsort_kernel<true>, binned.cu); ops without a fused store use fan_copy_kernel — a
//                         hand-written all-gather by peer stores — behind their kernels.  Two tiny flag-exchange
//                         kernels bracket the call: "my result array may be written" and "my stores have landed".
// NCCL is bound at run time (dlopen): the copy already loaded into the process (torch's) when there is one, so a
// torchrun rank never ends up with two NCCL runtimes; else libnccl.so.2 from the loader path.
#include <dlfcn.h>

#include <algorithm>
#include <new>

#include <nccl.h> // types and prototypes only; no link-time dependency

#include "fan.cuh"
#include "internal.h"

namespace sdslgpu
{

namespace
{

// ------------------------------------------------------------------------------------------------ NCCL at run time
struct NcclApi
{
    void * lib = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommInitAll) CommInitAll = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclBroadcast) Broadcast = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclGetVersion) GetVersion = nullptr;
    char where[256] = "";
};

NcclApi * nccl_api()
{
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        void * lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD); // already in the process (torch's bundled copy)?
        char const * how = "already loaded";
        if (!lib)
            if (char const * e = std::getenv("SDSLGPU_NCCL_LIB"))
            {
                lib = dlopen(e, RTLD_NOW | RTLD_GLOBAL);
                how = e;
            }
        if (!lib)
        {
            lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
            how = "libnccl.so.2";
        }
        if (!lib)
        {
            lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
            how = "libnccl.so";
        }
        if (!lib)
            return;
#define SG_SYM(name)                                                                                                   \
    api.name = reinterpret_cast<decltype(api.name)>(dlsym(lib, "nccl" #name));                                         \
    if (!api.name)                                                                                                     \
        return;
        SG_SYM(GetUniqueId)
        SG_SYM(CommInitRank)
        SG_SYM(CommInitAll)
        SG_SYM(CommDestroy)
        SG_SYM(AllGather)
        SG_SYM(Broadcast)
        SG_SYM(GroupStart)
        SG_SYM(GroupEnd)
        SG_SYM(GetErrorString)
        SG_SYM(GetVersion)
#undef SG_SYM
        snprintf(api.where, sizeof(api.where), "%s", how);
        api.lib = lib;
    });
    return api.lib ? &api : nullptr;
}

int nccl_fail(ncclResult_t r, char const * what, int line)
{
    NcclApi * a = nccl_api();
    set_error("NCCL error %d (%s) in %s at group.cu:%d", (int)r, a ? a->GetErrorString(r) : "?", what, line);
    return SDSLGPU_ECUDA;
}

#define SG_NCCL(expr)                                                                                                  \
    do                                                                                                                 \
    {                                                                                                                  \
        ncclResult_t r__ = (expr);                                                                                     \
        if (r__ != ncclSuccess)                                                                                        \
            return nccl_fail(r__, #expr, __LINE__);                                                                    \
    } while (0)

} // namespace

static constexpr int kMaxRanks = kMaxFan + 1;

struct GroupMember
{
    int device = 0;
    int rank = 0; // global rank of this local member
    ncclComm_t comm = nullptr;
    cudaStream_t stream = nullptr; // used when the caller passes no streams
    uint64_t * flags = nullptr;    // kMaxRanks arrival counters, written by the peers (symmetric memory)
    uint32_t * status = nullptr;   // device word set by a barrier that timed out
};

// one symmetric allocation: the same number of bytes on every member, every member can address every copy
struct SymBuf
{
    uint64_t bytes = 0;
    void * local[kMaxRanks] = {};            // [local member]
    void * peer[kMaxRanks][kMaxRanks] = {};  // [local member][global rank]: that rank's copy as seen from this member
    bool ipc = false;
};

} // namespace sdslgpu

struct sdslgpu_group
{
    int nranks = 0, nlocal = 0;
    bool loopback = false; // several members on one device (tests on a single GPU): no NCCL, peer stores only
    bool p2p = false;      // symmetric memory works (IPC handles could be opened / peer access enabled)
    sdslgpu::GroupMember m[sdslgpu::kMaxRanks];
    std::vector<sdslgpu::SymBuf> syms;
    sdslgpu::SymBuf stage; // SDSLGPU_GATHER_PACKED: per member one region per source rank, grown on demand
    uint64_t epoch = 0;
    std::mutex mu;
};

namespace sdslgpu
{

namespace
{

// ------------------------------------------------------------------------------------------------ device side
struct PeerFlags
{
    uint64_t * p[kMaxRanks];
};

__device__ __forceinline__ uint64_t global_ns()
{
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Flag exchange between the members of a group, one thread per peer: publish `epoch` in the peer's arrival counter for
// me (release at system scope: everything this stream stored before — including the peer stores of the previous
// kernel — is visible to whoever acquires the flag), then wait until that peer has published at least `epoch` in mine.
// Counters only grow, so nothing is ever reset.  A peer that never arrives (a failed rank) must not hang the GPU: after
// `timeout_ns` the kernel gives up and reports through *status.
__global__ void group_barrier_kernel(PeerFlags peers, uint64_t * mine, int me, int n, uint64_t epoch, uint64_t timeout_ns, uint32_t * status)
{
    int const r = (int)threadIdx.x;
    if (r >= n || r == me)
        return;
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(peers.p[r] + me), "l"(epoch) : "memory");
    uint64_t const t0 = global_ns();
    for (;;)
    {
        uint64_t v;
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine + r) : "memory");
        if (v >= epoch)
            break;
        if (global_ns() - t0 > timeout_ns)
        {
            atomicExch(status, 1u);
            break;
        }
        __nanosleep(200);
    }
}

// all-gather by peer stores for ops whose own kernels do not fan out: src[k] -> dst[r][k] for every peer r
__global__ void __launch_bounds__(kThreads) fan_copy_kernel(uint64_t const * __restrict__ src, Fan const fan, uint64_t n)
{
    uint64_t const stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride)
    {
        uint64_t const v = ld_stream_u64(src + k);
        for (uint32_t r = 0; r < fan.n; ++r)
            fan.dst[r][k] = v;
    }
}

// SDSLGPU_GATHER_PACKED for ops whose kernels do not fan out themselves: the shard's answers, packed, to every peer
__global__ void __launch_bounds__(kThreads) fan_pack_kernel(uint64_t const * __restrict__ src, Fan const fan, uint64_t n)
{
    uint64_t const stride = (uint64_t)gridDim.x * blockDim.x;
    uint32_t const lane = threadIdx.x & 31u;
    for (uint64_t p0 = (uint64_t)blockIdx.x * blockDim.x + (threadIdx.x - lane); p0 < n; p0 += stride)
    {
        uint64_t const p = p0 + lane;
        fan_store_packed(fan, p0, lane, n - p0 < 32 ? (uint32_t)(n - p0) : 32u, p < n ? ld_stream_u64(src + p) : 0ull);
    }
}

// the receiving side: widen the fields the other members stored into my staging regions into my result array.
// region r (source rank r != me) holds answers [r*s, (r+1)*s) as w-bit fields; all ones = SDSLGPU_NPOS
__global__ void __launch_bounds__(kThreads)
    fan_unpack_kernel(uint64_t const * __restrict__ stage, uint64_t region_words, uint32_t w, uint64_t s, int me, int nranks, uint64_t * __restrict__ out)
{
    uint64_t const total = (uint64_t)(nranks - 1) * s, stride = (uint64_t)gridDim.x * blockDim.x;
    uint64_t const mask = (1ull << w) - 1ull;
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride)
    {
        uint64_t r = e / s;
        uint64_t const i = e - r * s;
        r += r >= (uint64_t)me ? 1u : 0u;
        uint64_t const * reg = stage + r * region_words;
        uint64_t const pos = i * w, word = pos >> 6;
        uint32_t const off = (uint32_t)(pos & 63u);
        uint64_t v = reg[word] >> off;
        if (off + w > 64)
            v |= reg[word + 1] << (64 - off);
        v &= mask;
        st_stream_u64(out + r * s + i, v == mask ? SDSLGPU_NPOS : v);
    }
}

// ------------------------------------------------------------------------------------------------ host side
int enable_peer_access(sdslgpu_group * g)
{ // single process: every member's device maps every other member's memory
    for (int a = 0; a < g->nlocal; ++a)
        for (int b = 0; b < g->nlocal; ++b)
        {
            if (g->m[a].device == g->m[b].device)
                continue;
            int can = 0;
            SG_CUDA(cudaDeviceCanAccessPeer(&can, g->m[a].device, g->m[b].device));
            if (!can)
                return SDSLGPU_ENOTSUP;
            SG_CUDA(cudaSetDevice(g->m[a].device));
            cudaError_t e = cudaDeviceEnablePeerAccess(g->m[b].device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                return cuda_fail(e, "cudaDeviceEnablePeerAccess", __FILE__, __LINE__);
            cudaGetLastError();
        }
    return SDSLGPU_OK;
}

// collective: `bytes` on every member, zero-filled, and the map of everybody's copy
int sym_alloc(sdslgpu_group * g, uint64_t bytes, SymBuf & sb)
{
    sb = SymBuf();
    sb.bytes = bytes;
    for (int k = 0; k < g->nlocal; ++k)
    {
        SG_CUDA(cudaSetDevice(g->m[k].device));
        SG_CUDA(cudaMalloc(&sb.local[k], bytes ? bytes : 8));
        SG_CUDA(cudaMemset(sb.local[k], 0, bytes ? bytes : 8));
        SG_CUDA(cudaDeviceSynchronize());
    }
    if (g->nlocal == g->nranks)
    { // one process: unified addressing + peer access
        for (int k = 0; k < g->nlocal; ++k)
            for (int r = 0; r < g->nranks; ++r)
                sb.peer[k][r] = sb.local[r];
        return SDSLGPU_OK;
    }
    // one process per GPU: exchange CUDA IPC handles through the communicator
    NcclApi * nc = nccl_api();
    GroupMember & me = g->m[0];
    cudaIpcMemHandle_t mine;
    SG_CUDA(cudaIpcGetMemHandle(&mine, sb.local[0]));
    uint8_t * d_all = nullptr;
    size_t const hb = sizeof(cudaIpcMemHandle_t);
    SG_CUDA(cudaMalloc(reinterpret_cast<void **>(&d_all), hb * g->nranks));
    SG_CUDA(cudaMemcpy(d_all + hb * me.rank, &mine, hb, cudaMemcpyHostToDevice));
    ncclResult_t nr = nc->AllGather(d_all + hb * me.rank, d_all, hb, ncclChar, me.comm, me.stream);
    cudaError_t ce = cudaStreamSynchronize(me.stream);
    std::vector<cudaIpcMemHandle_t> all(g->nranks);
    if (nr == ncclSuccess && ce == cudaSuccess)
        ce = cudaMemcpy(all.data(), d_all, hb * g->nranks, cudaMemcpyDeviceToHost);
    cudaFree(d_all);
    if (nr != ncclSuccess)
        return nccl_fail(nr, "ncclAllGather (IPC handles)", __LINE__);
    if (ce != cudaSuccess)
        return cuda_fail(ce, "IPC handle exchange", __FILE__, __LINE__);
    sb.ipc = true;
    for (int r = 0; r < g->nranks; ++r)
    {
        if (r == me.rank)
        {
            sb.peer[0][r] = sb.local[0];
            continue;
        }
        cudaError_t e = cudaIpcOpenMemHandle(&sb.peer[0][r], all[r], cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess)
            return cuda_fail(e, "cudaIpcOpenMemHandle", __FILE__, __LINE__);
    }
    return SDSLGPU_OK;
}

void sym_free(sdslgpu_group * g, SymBuf & sb)
{
    for (int k = 0; k < g->nlocal; ++k)
    {
        cudaSetDevice(g->m[k].device);
        if (sb.ipc)
            for (int r = 0; r < g->nranks; ++r)
                if (r != g->m[k].rank && sb.peer[k][r])
                    cudaIpcCloseMemHandle(sb.peer[k][r]);
        if (sb.local[k])
            cudaFree(sb.local[k]);
    }
    cudaGetLastError();
    sb = SymBuf();
}

// a collective no-op on every member's stream: nobody passes before everybody has arrived (used before frees)
int nccl_barrier(sdslgpu_group * g)
{
    if (g->loopback || g->nranks == 1)
    {
        for (int k = 0; k < g->nlocal; ++k)
        {
            SG_CUDA(cudaSetDevice(g->m[k].device));
            SG_CUDA(cudaDeviceSynchronize());
        }
        return SDSLGPU_OK;
    }
    NcclApi * nc = nccl_api();
    std::vector<uint8_t *> tmp(g->nlocal, nullptr);
    for (int k = 0; k < g->nlocal; ++k)
    {
        SG_CUDA(cudaSetDevice(g->m[k].device));
        SG_CUDA(cudaDeviceSynchronize());
        SG_CUDA(cudaMalloc(reinterpret_cast<void **>(&tmp[k]), 8 * g->nranks));
    }
    SG_NCCL(nc->GroupStart());
    for (int k = 0; k < g->nlocal; ++k)
        nc->AllGather(tmp[k] + 8 * g->m[k].rank, tmp[k], 8, ncclChar, g->m[k].comm, g->m[k].stream);
    SG_NCCL(nc->GroupEnd());
    for (int k = 0; k < g->nlocal; ++k)
    {
        SG_CUDA(cudaSetDevice(g->m[k].device));
        SG_CUDA(cudaStreamSynchronize(g->m[k].stream));
        cudaFree(tmp[k]);
    }
    return SDSLGPU_OK;
}

int finish_create(sdslgpu_group * g)
{
    for (int k = 0; k < g->nlocal; ++k)
    {
        SG_CUDA(cudaSetDevice(g->m[k].device));
        cudaFuncAttributes fa; // loads the modules of the exchange / copy kernels now, not behind a spinning kernel
        SG_CUDA(cudaFuncGetAttributes(&fa, group_barrier_kernel));
        SG_CUDA(cudaFuncGetAttributes(&fa, fan_copy_kernel));
        SG_CUDA(cudaFuncGetAttributes(&fa, fan_pack_kernel));
        SG_CUDA(cudaFuncGetAttributes(&fa, fan_unpack_kernel));
        SG_CUDA(cudaStreamCreateWithFlags(&g->m[k].stream, cudaStreamNonBlocking));
        SG_CUDA(cudaMalloc(reinterpret_cast<void **>(&g->m[k].status), 256));
        SG_CUDA(cudaMemset(g->m[k].status, 0, 256));
    }
    g->p2p = false;
    if (g->nranks > 1)
    {
        bool ok = true;
        if (g->nlocal == g->nranks && !g->loopback)
            ok = enable_peer_access(g) == SDSLGPU_OK;
        if (ok && std::getenv("SDSLGPU_GROUP_NO_P2P") == nullptr)
        {
            SymBuf sb;
            if (sym_alloc(g, 8 * kMaxRanks, sb) == SDSLGPU_OK)
            {
                g->syms.push_back(sb);
                for (int k = 0; k < g->nlocal; ++k)
                    g->m[k].flags = static_cast<uint64_t *>(sb.local[k]);
                g->p2p = true;
            }
            else
                sym_free(g, sb); // peers cannot map each other's memory here: the NCCL gather still works
        }
    }
    return SDSLGPU_OK;
}

// the symmetric allocation that contains [p, p + bytes) on local member k, or nullptr
SymBuf const * find_sym(sdslgpu_group const * g, int k, void const * p, uint64_t bytes, uint64_t * offset)
{
    for (SymBuf const & sb : g->syms)
    {
        uint8_t const * b = static_cast<uint8_t const *>(sb.local[k]);
        uint8_t const * q = static_cast<uint8_t const *>(p);
        if (q >= b && q + bytes <= b + sb.bytes)
        {
            *offset = (uint64_t)(q - b);
            return &sb;
        }
    }
    return nullptr;
}

int launch_barrier(sdslgpu_group * g, int k, uint64_t epoch, cudaStream_t s)
{
    PeerFlags pf;
    SymBuf const & fb = g->syms[0];
    for (int r = 0; r < g->nranks; ++r)
        pf.p[r] = static_cast<uint64_t *>(fb.peer[k][r]);
    uint64_t timeout_ns = 20ull * 1000 * 1000 * 1000;
    if (char const * e = std::getenv("SDSLGPU_GROUP_TIMEOUT_MS"))
        timeout_ns = (uint64_t)std::atoll(e) * 1000 * 1000;
    group_barrier_kernel<<<1, kMaxRanks, 0, s>>>(pf, g->m[k].flags, g->m[k].rank, g->nranks, epoch, timeout_ns, g->m[k].status);
    SG_CUDA(cudaGetLastError());
    return SDSLGPU_OK;
}

// The sharded call.  shard(k, first, count, out, stream, fan, &fanned) launches member k's kernels for queries
// [first, first + count) writing out[0 .. count) (a pointer already offset to `first`).
// bits that hold every possible answer of an op (answers <= max_answer) with the all-ones pattern left free for
// SDSLGPU_NPOS, rounded up to what fan_store_packed handles (even, 22 .. 62); 0: the op cannot travel packed
uint32_t pack_width_for(uint64_t max_answer)
{
    uint32_t w = 1;
    while (w < 64 && ((1ull << w) - 1ull) <= max_answer)
        ++w;
    w += w & 1u;
    if (w < kPackMinWidth)
        w = kPackMinWidth;
    return w <= kPackMaxWidth ? w : 0u;
}

template <class Shard>
int group_run(sdslgpu_group * g, uint64_t n, uint64_t * const * out, int gather, void * const * streams, uint32_t pack_width, Shard shard)
{
    if (!g || !out)
    {
        set_error("group call: null argument");
        return SDSLGPU_EINVAL;
    }
    std::lock_guard<std::mutex> lock(g->mu);
    int prev_dev = -1;
    cudaGetDevice(&prev_dev);
    struct Restore
    {
        int d;
        ~Restore()
        {
            if (d >= 0)
                cudaSetDevice(d);
        }
    } restore{prev_dev};
    uint64_t const s = n / (uint64_t)g->nranks, covered = s * (uint64_t)g->nranks;
    // which gather: FUSED needs every member's result array inside one symmetric allocation at the same offset
    bool fused_ok = g->p2p && g->nranks > 1;
    uint64_t off0 = 0;
    SymBuf const * sb = nullptr;
    if (fused_ok && gather != SDSLGPU_GATHER_NONE && gather != SDSLGPU_GATHER_NCCL)
        for (int k = 0; k < g->nlocal; ++k)
        {
            uint64_t off = 0;
            SymBuf const * f = find_sym(g, k, out[k], n * 8, &off);
            if (!f || (sb && (f != sb || off != off0)))
            {
                fused_ok = false;
                break;
            }
            sb = f;
            off0 = off;
        }
    bool const packed_ok = g->p2p && g->nranks > 1 && pack_width != 0 && std::getenv("SDSLGPU_GROUP_NO_PACKED") == nullptr;
    int mode = gather;
    if (g->nranks == 1)
        mode = SDSLGPU_GATHER_NONE;
    else if (gather == SDSLGPU_GATHER_AUTO)
        mode = fused_ok ? SDSLGPU_GATHER_FUSED : packed_ok ? SDSLGPU_GATHER_PACKED : SDSLGPU_GATHER_NCCL; // measured order
    if (mode == SDSLGPU_GATHER_PACKED && !packed_ok)
    {
        set_error("SDSLGPU_GATHER_PACKED needs peer-mapped memory between the members (and answers of at most 62 bits)");
        return SDSLGPU_EINVAL;
    }
    if (mode == SDSLGPU_GATHER_FUSED && !fused_ok)
    {
        set_error("SDSLGPU_GATHER_FUSED needs result arrays from sdslgpu_group_alloc (same offset on every member) and peer-mapped memory");
        return SDSLGPU_EINVAL;
    }
    if (mode == SDSLGPU_GATHER_NCCL && (g->loopback || !nccl_api()))
    {
        set_error("SDSLGPU_GATHER_NCCL: no NCCL communicator in this group (%s)", g->loopback ? "loopback group on one device" : "libnccl.so.2 not found");
        return SDSLGPU_ENOTSUP;
    }
    bool const peer_stores = mode == SDSLGPU_GATHER_FUSED || mode == SDSLGPU_GATHER_PACKED;
    uint64_t const region_words = fan_region_words(s, pack_width) + 1; // + 1: the receiver reads two words per field
    if (mode == SDSLGPU_GATHER_PACKED && g->stage.bytes < (uint64_t)g->nranks * region_words * 8)
    { // first packed call (or a larger batch): everybody re-allocates its staging regions — collective, like the call itself
        SG_TRY(nccl_barrier(g));
        sym_free(g, g->stage);
        SG_TRY(sym_alloc(g, (uint64_t)g->nranks * region_words * 8 + ((uint64_t)g->nranks * region_words * 8) / 4, g->stage));
    }
    uint64_t const ep_in = g->epoch + 1, ep_out = g->epoch + 2;
    if (peer_stores)
        g->epoch += 2;
    std::vector<cudaStream_t> st(g->nlocal);
    for (int k = 0; k < g->nlocal; ++k)
        st[k] = streams ? static_cast<cudaStream_t>(streams[k]) : g->m[k].stream; // a NULL entry is the legacy default stream
    bool const own_streams = streams == nullptr;
    // Three passes over the local members, never a kernel launch of member k behind a flag exchange that waits for a
    // member this same host thread has not launched yet: the first launch of a kernel loads its module, which can
    // synchronise the context, and a spinning exchange kernel would then wait for a launch that cannot be issued.
    std::vector<Fan> fans(g->nlocal);
    if (peer_stores)
        for (int k = 0; k < g->nlocal; ++k)
        {
            GroupMember & me = g->m[k];
            SG_CUDA(cudaSetDevice(me.device));
            for (int r = 0; r < g->nranks; ++r)
            {
                if (r == me.rank)
                    continue;
                if (mode == SDSLGPU_GATHER_PACKED) // my region of member r's staging buffer
                    fans[k].dst[fans[k].n++] = static_cast<uint64_t *>(g->stage.peer[k][r]) + (uint64_t)me.rank * region_words;
                else
                    fans[k].dst[fans[k].n++] = reinterpret_cast<uint64_t *>(static_cast<uint8_t *>(sb->peer[k][r]) + off0) + (uint64_t)me.rank * s;
            }
            fans[k].width = mode == SDSLGPU_GATHER_PACKED ? pack_width : 0u;
            SG_TRY(launch_barrier(g, k, ep_in, st[k])); // every member has reached this call: its arrays may be written
        }
    for (int k = 0; k < g->nlocal; ++k)
    {
        GroupMember & me = g->m[k];
        SG_CUDA(cudaSetDevice(me.device));
        uint64_t const first = (uint64_t)me.rank * s;
        Fan const & fan = fans[k];
        if (s)
        {
            bool fanned = false;
            SG_TRY(shard(k, first, s, out[k] + first, st[k], fan.n ? &fan : nullptr, &fanned));
            if (fan.n && !fanned)
            {
                if (fan.width)
                    fan_pack_kernel<<<grid_for(s), kThreads, 0, st[k]>>>(out[k] + first, fan, s);
                else
                    fan_copy_kernel<<<grid_for(s), kThreads, 0, st[k]>>>(out[k] + first, fan, s);
                SG_CUDA(cudaGetLastError());
            }
        }
        if (covered < n) // the < nranks queries that do not divide evenly are answered by everybody
        {
            bool fanned = false;
            SG_TRY(shard(k, covered, n - covered, out[k] + covered, st[k], nullptr, &fanned));
        }
    }
    if (peer_stores)
        for (int k = 0; k < g->nlocal; ++k)
        {
            SG_CUDA(cudaSetDevice(g->m[k].device));
            SG_TRY(launch_barrier(g, k, ep_out, st[k])); // everybody's stores into my arrays have landed
            if (mode == SDSLGPU_GATHER_PACKED && s)
            {
                fan_unpack_kernel<<<grid_for((uint64_t)(g->nranks - 1) * s), kThreads, 0, st[k]>>>(static_cast<uint64_t const *>(g->stage.local[k]), region_words,
                                                                                                 pack_width, s, g->m[k].rank, g->nranks, out[k]);
                SG_CUDA(cudaGetLastError());
            }
        }
    if (mode == SDSLGPU_GATHER_NCCL && s)
    {
        NcclApi * nc = nccl_api();
        SG_NCCL(nc->GroupStart());
        for (int k = 0; k < g->nlocal; ++k)
        {
            ncclResult_t r = nc->AllGather(out[k] + (uint64_t)g->m[k].rank * s, out[k], s, ncclUint64, g->m[k].comm, st[k]);
            if (r != ncclSuccess)
            {
                nc->GroupEnd();
                return nccl_fail(r, "ncclAllGather (results)", __LINE__);
            }
        }
        SG_NCCL(nc->GroupEnd());
    }
    if (own_streams)
        for (int k = 0; k < g->nlocal; ++k)
        {
            SG_CUDA(cudaSetDevice(g->m[k].device));
            SG_CUDA(cudaStreamSynchronize(st[k]));
            if (peer_stores)
            {
                uint32_t bad = 0;
                SG_CUDA(cudaMemcpy(&bad, g->m[k].status, 4, cudaMemcpyDeviceToHost));
                if (bad)
                {
                    set_error("group call: a member did not arrive at the flag exchange in time (member %d waited)", g->m[k].rank);
                    return SDSLGPU_ECUDA;
                }
            }
        }
    return SDSLGPU_OK;
}

int check_members(sdslgpu_group const * g, sdslgpu_handle const * const * h, int kind_a, int kind_b, char const * who)
{
    if (!g || !h)
    {
        set_error("%s: null argument", who);
        return SDSLGPU_EINVAL;
    }
    for (int k = 0; k < g->nlocal; ++k)
    {
        if (!h[k] || (h[k]->kind != kind_a && h[k]->kind != kind_b))
        {
            set_error("%s: member %d has no handle of the right kind", who, k);
            return SDSLGPU_EINVAL;
        }
        if (h[k]->device != g->m[k].device)
        {
            set_error("%s: handle %d lives on device %d, the group member on device %d", who, k, h[k]->device, g->m[k].device);
            return SDSLGPU_EINVAL;
        }
    }
    return SDSLGPU_OK;
}

} // namespace
} // namespace sdslgpu

using namespace sdslgpu;

extern "C"
{

    int sdslgpu_group_unique_id(void * id)
    {
        if (!id)
            return SDSLGPU_EINVAL;
        NcclApi * nc = nccl_api();
        if (!nc)
        {
            set_error("sdslgpu_group_unique_id: libnccl.so.2 not found (set SDSLGPU_NCCL_LIB)");
            return SDSLGPU_ENOTSUP;
        }
        static_assert(sizeof(ncclUniqueId) == SDSLGPU_UNIQUE_ID_BYTES, "ncclUniqueId is 128 bytes");
        ncclUniqueId u;
        SG_NCCL(nc->GetUniqueId(&u));
        std::memcpy(id, &u, sizeof(u));
        return SDSLGPU_OK;
    }

    int sdslgpu_group_create_rank(const void * id, int nranks, int rank, int device, sdslgpu_group ** out)
    {
        if (!id || !out || nranks < 1 || nranks > kMaxRanks || rank < 0 || rank >= nranks)
        {
            set_error("sdslgpu_group_create_rank: bad argument (1 <= nranks <= %d)", kMaxRanks);
            return SDSLGPU_EINVAL;
        }
        *out = nullptr;
        int ndev = 0;
        SG_CUDA(cudaGetDeviceCount(&ndev));
        if (device < 0 || device >= ndev)
        {
            set_error("sdslgpu_group_create_rank: device %d out of range (%d devices); there is no CPU fallback", device, ndev);
            return SDSLGPU_ECUDA;
        }
        NcclApi * nc = nccl_api();
        if (!nc)
        {
            set_error("sdslgpu_group_create_rank: libnccl.so.2 not found (set SDSLGPU_NCCL_LIB)");
            return SDSLGPU_ENOTSUP;
        }
        DeviceGuard dg(device);
        sdslgpu_group * g = new (std::nothrow) sdslgpu_group;
        if (!g)
            return SDSLGPU_ENOMEM;
        g->nranks = nranks;
        g->nlocal = 1;
        g->m[0].device = device;
        g->m[0].rank = rank;
        ncclUniqueId u;
        std::memcpy(&u, id, sizeof(u));
        ncclResult_t r = nc->CommInitRank(&g->m[0].comm, nranks, u, rank);
        if (r != ncclSuccess)
        {
            delete g;
            return nccl_fail(r, "ncclCommInitRank", __LINE__);
        }
        int st = finish_create(g);
        if (st != SDSLGPU_OK)
        {
            sdslgpu_group_free(g);
            return st;
        }
        *out = g;
        return SDSLGPU_OK;
    }

    int sdslgpu_group_create(const int * devices, int ndev, sdslgpu_group ** out)
    {
        if (!devices || !out || ndev < 1 || ndev > kMaxRanks)
        {
            set_error("sdslgpu_group_create: bad argument (1 <= ndev <= %d)", kMaxRanks);
            return SDSLGPU_EINVAL;
        }
        *out = nullptr;
        int have = 0;
        SG_CUDA(cudaGetDeviceCount(&have));
        bool dup = false;
        for (int k = 0; k < ndev; ++k)
        {
            if (devices[k] < 0 || devices[k] >= have)
            {
                set_error("sdslgpu_group_create: device %d out of range (%d devices); there is no CPU fallback", devices[k], have);
                return SDSLGPU_ECUDA;
            }
            for (int j = 0; j < k; ++j)
                dup |= devices[j] == devices[k];
        }
        int prev = -1;
        cudaGetDevice(&prev);
        sdslgpu_group * g = new (std::nothrow) sdslgpu_group;
        if (!g)
            return SDSLGPU_ENOMEM;
        g->nranks = g->nlocal = ndev;
        g->loopback = dup;
        for (int k = 0; k < ndev; ++k)
        {
            g->m[k].device = devices[k];
            g->m[k].rank = k;
        }
        int st = SDSLGPU_OK;
        if (!dup && ndev > 1)
        {
            NcclApi * nc = nccl_api();
            if (!nc)
            {
                set_error("sdslgpu_group_create: libnccl.so.2 not found (set SDSLGPU_NCCL_LIB)");
                st = SDSLGPU_ENOTSUP;
            }
            else
            {
                ncclComm_t comms[kMaxRanks];
                ncclResult_t r = nc->CommInitAll(comms, ndev, devices);
                if (r != ncclSuccess)
                    st = nccl_fail(r, "ncclCommInitAll", __LINE__);
                else
                    for (int k = 0; k < ndev; ++k)
                        g->m[k].comm = comms[k];
            }
        }
        if (st == SDSLGPU_OK)
            st = finish_create(g);
        if (prev >= 0)
            cudaSetDevice(prev);
        if (st != SDSLGPU_OK)
        {
            sdslgpu_group_free(g);
            return st;
        }
        *out = g;
        return SDSLGPU_OK;
    }

    int sdslgpu_group_free(sdslgpu_group * g)
    {
        if (!g)
            return SDSLGPU_OK;
        int prev = -1;
        cudaGetDevice(&prev);
        nccl_barrier(g); // nobody unmaps memory a peer may still be storing into
        for (SymBuf & sb : g->syms)
            sym_free(g, sb);
        g->syms.clear();
        sym_free(g, g->stage);
        NcclApi * nc = nccl_api();
        for (int k = 0; k < g->nlocal; ++k)
        {
            cudaSetDevice(g->m[k].device);
            if (g->m[k].comm && nc)
                nc->CommDestroy(g->m[k].comm);
            if (g->m[k].stream)
                cudaStreamDestroy(g->m[k].stream);
            if (g->m[k].status)
                cudaFree(g->m[k].status);
        }
        cudaGetLastError();
        if (prev >= 0)
            cudaSetDevice(prev);
        delete g;
        return SDSLGPU_OK;
    }

    int sdslgpu_group_info(const sdslgpu_group * g, int * nranks, int * nlocal, int * first_rank, int * fused_possible)
    {
        if (!g)
            return SDSLGPU_EINVAL;
        if (nranks)
            *nranks = g->nranks;
        if (nlocal)
            *nlocal = g->nlocal;
        if (first_rank)
            *first_rank = g->m[0].rank;
        if (fused_possible)
            *fused_possible = g->p2p ? 1 : 0;
        return SDSLGPU_OK;
    }

    int sdslgpu_group_alloc(sdslgpu_group * g, uint64_t bytes, void ** ptrs)
    {
        if (!g || !ptrs)
            return SDSLGPU_EINVAL;
        std::lock_guard<std::mutex> lock(g->mu);
        int prev = -1;
        cudaGetDevice(&prev);
        int st = SDSLGPU_OK;
        if (g->nranks > 1 && !g->p2p)
        { // no peer mapping on this box: plain device memory, the NCCL gather works on it
            for (int k = 0; k < g->nlocal && st == SDSLGPU_OK; ++k)
            {
                cudaSetDevice(g->m[k].device);
                cudaError_t e = cudaMalloc(&ptrs[k], bytes ? bytes : 8);
                if (e != cudaSuccess)
                    st = cuda_fail(e, "cudaMalloc", __FILE__, __LINE__);
            }
            SymBuf sb;
            sb.bytes = bytes;
            for (int k = 0; k < g->nlocal; ++k)
                sb.local[k] = ptrs[k];
            g->syms.push_back(sb); // remembered so that sdslgpu_group_release frees it
        }
        else
        {
            SymBuf sb;
            st = sym_alloc(g, bytes, sb);
            if (st == SDSLGPU_OK)
            {
                for (int k = 0; k < g->nlocal; ++k)
                    ptrs[k] = sb.local[k];
                g->syms.push_back(sb);
            }
            else
                sym_free(g, sb);
        }
        if (prev >= 0)
            cudaSetDevice(prev);
        return st;
    }

    int sdslgpu_group_release(sdslgpu_group * g, void * const * ptrs)
    {
        if (!g || !ptrs)
            return SDSLGPU_EINVAL;
        std::lock_guard<std::mutex> lock(g->mu);
        int prev = -1;
        cudaGetDevice(&prev);
        int st = SDSLGPU_EINVAL;
        for (size_t i = g->p2p ? 1 : 0; i < g->syms.size(); ++i) // with peer mapping, entry 0 is the group's own flag buffer
            if (g->syms[i].local[0] == ptrs[0])
            {
                st = nccl_barrier(g);
                sym_free(g, g->syms[i]);
                g->syms.erase(g->syms.begin() + (long)i);
                break;
            }
        if (st == SDSLGPU_EINVAL)
            set_error("sdslgpu_group_release: not a pointer from sdslgpu_group_alloc");
        if (prev >= 0)
            cudaSetDevice(prev);
        return st;
    }

    // ---------------------------------------------------------------------------------------- replicate
    static int replicate_what(sdslgpu_handle const * h)
    {
        return h->kind == SDSLGPU_KIND_SD ? 1 : 0; // the complete sd_vector bytes; everything else: what 0
    }

    int sdslgpu_group_replicate(sdslgpu_group * g, const sdslgpu_handle * src, int root, sdslgpu_handle ** out)
    {
        if (!g || !out || root < 0 || root >= g->nranks)
        {
            set_error("sdslgpu_group_replicate: bad argument");
            return SDSLGPU_EINVAL;
        }
        bool root_here = false;
        for (int k = 0; k < g->nlocal; ++k)
            root_here |= g->m[k].rank == root;
        if (root_here && !src)
        {
            set_error("sdslgpu_group_replicate: the root rank must pass the handle to replicate");
            return SDSLGPU_EINVAL;
        }
        // header: kind, flags, sa_dens, isa_dens, order, nbytes — then the reference-format blob (sdslgpu_serialize)
        uint64_t hdr[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        std::vector<uint8_t> blob;
        int root_status = SDSLGPU_OK;
        if (root_here)
        { // a failure here must still reach the other ranks (hdr[6] stays 0): they are about to wait in ncclBroadcast
            uint64_t nb = 0;
            int what = replicate_what(src);
            root_status = sdslgpu_serialize(src, what, nullptr, 0, &nb);
            if (root_status == SDSLGPU_OK)
            {
                blob.resize(nb ? nb : 1);
                root_status = sdslgpu_serialize(src, what, blob.data(), nb, &nb);
                blob.resize(nb);
            }
            if (root_status == SDSLGPU_OK)
            {
                hdr[0] = (uint64_t)src->kind;
                hdr[1] = src->flags & ~(uint32_t)SDSLGPU_F_V5_SCAN;
                hdr[2] = src->kind == SDSLGPU_KIND_CSA_WT ? src->csa.sa_dens : 0;
                hdr[3] = src->kind == SDSLGPU_KIND_CSA_WT ? src->csa.isa_dens : 0;
                hdr[4] = (uint64_t)src->order;
                hdr[5] = nb;
                hdr[6] = 1; // valid
            }
        }
        int prev = -1;
        cudaGetDevice(&prev);
        struct Restore
        {
            int d;
            ~Restore()
            {
                if (d >= 0)
                    cudaSetDevice(d);
            }
        } restore{prev};
        if (g->nlocal < g->nranks)
        { // one process per GPU: header and blob travel through ncclBroadcast
            NcclApi * nc = nccl_api();
            GroupMember & me = g->m[0];
            SG_CUDA(cudaSetDevice(me.device));
            uint64_t * d_hdr = nullptr;
            SG_CUDA(cudaMalloc(reinterpret_cast<void **>(&d_hdr), sizeof(hdr)));
            SG_CUDA(cudaMemcpy(d_hdr, hdr, sizeof(hdr), cudaMemcpyHostToDevice));
            ncclResult_t r = nc->Broadcast(d_hdr, d_hdr, sizeof(hdr), ncclChar, root, me.comm, me.stream);
            cudaError_t e = cudaStreamSynchronize(me.stream);
            if (r == ncclSuccess && e == cudaSuccess)
                e = cudaMemcpy(hdr, d_hdr, sizeof(hdr), cudaMemcpyDeviceToHost);
            cudaFree(d_hdr);
            if (r != ncclSuccess)
                return nccl_fail(r, "ncclBroadcast (header)", __LINE__);
            if (e != cudaSuccess)
                return cuda_fail(e, "replicate header", __FILE__, __LINE__);
            if (hdr[6] != 1)
            {
                if (!root_here)
                    set_error("sdslgpu_group_replicate: the root rank failed to serialise its handle");
                return root_here ? root_status : SDSLGPU_EINVAL;
            }
            uint64_t const nb = hdr[5];
            uint8_t * d_blob = nullptr;
            SG_CUDA(cudaMalloc(reinterpret_cast<void **>(&d_blob), nb ? nb : 8));
            if (root_here)
                SG_CUDA(cudaMemcpy(d_blob, blob.data(), nb, cudaMemcpyHostToDevice));
            r = nc->Broadcast(d_blob, d_blob, nb, ncclChar, root, me.comm, me.stream);
            e = cudaStreamSynchronize(me.stream);
            if (r == ncclSuccess && e == cudaSuccess && !root_here)
            {
                blob.resize(nb ? nb : 1);
                e = cudaMemcpy(blob.data(), d_blob, nb, cudaMemcpyDeviceToHost);
            }
            cudaFree(d_blob);
            if (r != ncclSuccess)
                return nccl_fail(r, "ncclBroadcast (blob)", __LINE__);
            if (e != cudaSuccess)
                return cuda_fail(e, "replicate blob", __FILE__, __LINE__);
        }
        if (root_status != SDSLGPU_OK)
            return root_status; // one process driving all devices: nobody is waiting
        for (int k = 0; k < g->nlocal; ++k)
        {
            out[k] = nullptr;
            sdslgpu_handle * h = nullptr;
            SG_TRY(sdslgpu_load_sdsl_ex(blob.data(), hdr[5], (int)hdr[0], g->m[k].device, (uint32_t)hdr[1], (uint32_t)hdr[2], (uint32_t)hdr[3], nullptr, &h));
            if ((int)hdr[0] == SDSLGPU_KIND_BV || (int)hdr[0] == SDSLGPU_KIND_SD || (int)hdr[0] == SDSLGPU_KIND_RRR63)
                sdslgpu_set_batch_order(h, (int)hdr[4]);
            out[k] = h;
        }
        return SDSLGPU_OK;
    }

    // ---------------------------------------------------------------------------------------- sharded queries
    // bit-vector handles of any kind: plain (fan-out fused into its kernels), rrr_vector<63>, sd_vector<>
    static int check_bitvector_members(sdslgpu_group const * g, sdslgpu_handle const * const * h, int b, char const * who)
    {
        if (!g || !h)
        {
            set_error("%s: null argument", who);
            return SDSLGPU_EINVAL;
        }
        if (b != 0 && b != 1)
        {
            set_error("%s: b must be 0 or 1", who);
            return SDSLGPU_EINVAL;
        }
        for (int k = 0; k < g->nlocal; ++k)
        {
            if (!h[k] || (h[k]->kind != SDSLGPU_KIND_BV && h[k]->kind != SDSLGPU_KIND_RRR63 && h[k]->kind != SDSLGPU_KIND_SD) || h[k]->kind != h[0]->kind)
            {
                set_error("%s: member %d has no bit-vector handle (plain, rrr, sd; the same kind on every member)", who, k);
                return SDSLGPU_EINVAL;
            }
            if (h[k]->device != g->m[k].device)
            {
                set_error("%s: handle %d lives on device %d, the group member on device %d", who, k, h[k]->device, g->m[k].device);
                return SDSLGPU_EINVAL;
            }
        }
        return SDSLGPU_OK;
    }

    int sdslgpu_group_rank(sdslgpu_group * g, const sdslgpu_handle * const * h, int b, const uint64_t * const * idx, uint64_t n, uint64_t * const * out,
                           int gather, void * const * streams)
    {
        SG_TRY(check_bitvector_members(g, h, b, "sdslgpu_group_rank"));
        uint64_t size = 0;
        SG_TRY(sdslgpu_size(h[0], &size));
        return group_run(g, n, out, gather, streams, pack_width_for(size), [&](int k, uint64_t first, uint64_t cnt, uint64_t * o, cudaStream_t s, Fan const * fan, bool * fanned) {
            *fanned = false;
            if (h[k]->kind == SDSLGPU_KIND_RRR63)
                return rrr_rank_device(h[k], b, idx[k] + first, cnt, o, s);
            if (h[k]->kind == SDSLGPU_KIND_SD)
                return sd_rank_device(h[k], b, idx[k] + first, cnt, o, s);
            return bv_rank_device(h[k]->bv, h[k]->flags, b, idx[k] + first, cnt, o, s, fan, fanned);
        });
    }

    int sdslgpu_group_select(sdslgpu_group * g, const sdslgpu_handle * const * h, int b, const uint64_t * const * i, uint64_t n, uint64_t * const * out,
                             int gather, void * const * streams)
    {
        SG_TRY(check_bitvector_members(g, h, b, "sdslgpu_group_select"));
        uint64_t size = 0;
        SG_TRY(sdslgpu_size(h[0], &size));
        return group_run(g, n, out, gather, streams, pack_width_for(size), [&](int k, uint64_t first, uint64_t cnt, uint64_t * o, cudaStream_t s, Fan const * fan, bool * fanned) {
            *fanned = false;
            if (h[k]->kind == SDSLGPU_KIND_RRR63)
                return rrr_select_device(h[k], b, i[k] + first, cnt, o, s);
            if (h[k]->kind == SDSLGPU_KIND_SD)
                return sd_select_device(h[k], b, i[k] + first, cnt, o, s);
            SG_TRY(bv_ensure_select_sectors(h[k], b, cnt));
            return bv_select_device(h[k]->bv, b, i[k] + first, cnt, o, s, fan, fanned);
        });
    }

    int sdslgpu_group_wt_rank(sdslgpu_group * g, const sdslgpu_handle * const * h, const uint64_t * const * i, const uint8_t * const * c, uint64_t n,
                              uint64_t * const * out, int gather, void * const * streams)
    {
        SG_TRY(check_members(g, h, SDSLGPU_KIND_WT_HUFF, SDSLGPU_KIND_CSA_WT, "sdslgpu_group_wt_rank"));
        uint64_t size = 0;
        SG_TRY(sdslgpu_size(h[0], &size));
        return group_run(g, n, out, gather, streams, pack_width_for(size), [&](int k, uint64_t first, uint64_t cnt, uint64_t * o, cudaStream_t s, Fan const *, bool * fanned) {
            *fanned = false;
            return wt_rank_device(h[k], i[k] + first, c[k] + first, cnt, o, s);
        });
    }

    int sdslgpu_group_fm_count(sdslgpu_group * g, const sdslgpu_handle * const * h, const uint8_t * const * pats, const uint64_t * const * pat_off, uint64_t n,
                               uint64_t * const * cnt_out, int gather, void * const * streams)
    {
        SG_TRY(check_members(g, h, SDSLGPU_KIND_CSA_WT, SDSLGPU_KIND_CSA_WT, "sdslgpu_group_fm_count"));
        uint64_t size = 0;
        SG_TRY(sdslgpu_size(h[0], &size));
        return group_run(g, n, cnt_out, gather, streams, pack_width_for(size), [&](int k, uint64_t first, uint64_t cnt, uint64_t * o, cudaStream_t s, Fan const *, bool * fanned) {
            *fanned = false;
            return fm_count_device(h[k], pats[k], pat_off[k] + first, cnt, o, nullptr, s);
        });
    }

} // extern "C"
