// api.cu — the extern "C" boundary declared in include/sdslgpu.h: handle lifetime, pointer
// classification, the chunked host-buffer pipeline, and dispatch to the per-kind device launchers.
#include <cstdarg>
#include <new>

#include "internal.h"

namespace sdslgpu
{

static thread_local char g_err[512] = "";

void set_error(char const * fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, char const * what, char const * file, int line)
{
    set_error("CUDA error %d (%s) at %s:%d in %s", (int)e, cudaGetErrorString(e), file, line, what);
    cudaGetLastError(); // clear the sticky-free error state
    return e == cudaErrorMemoryAllocation ? SDSLGPU_ENOMEM : SDSLGPU_ECUDA;
}

int sm_count()
{
    static int cached[64] = {0}; // per device ordinal; a benign race writes the same value twice
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64)
        return 148;
    if (cached[dev] == 0)
    {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
        {
            cudaGetLastError();
            n = 148;
        }
        cached[dev] = n;
    }
    return cached[dev];
}

int DevicePool::alloc(void ** p, uint64_t n)
{
    *p = nullptr;
    if (n == 0)
        n = 8;
    n = (n + 255) & ~255ull;
    cudaError_t e = cudaMalloc(p, n);
    if (e != cudaSuccess)
        return cuda_fail(e, "cudaMalloc", __FILE__, __LINE__);
    ptrs.push_back(*p);
    sizes_.push_back(n);
    bytes += n;
    return SDSLGPU_OK;
}

void DevicePool::release(void * p)
{
    for (size_t k = 0; k < ptrs.size(); ++k)
        if (ptrs[k] == p)
        {
            cudaFree(p);
            bytes -= sizes_[k];
            ptrs.erase(ptrs.begin() + k);
            sizes_.erase(sizes_.begin() + k);
            return;
        }
}

void DevicePool::release_all()
{
    for (void * p : ptrs)
        cudaFree(p);
    ptrs.clear();
    sizes_.clear();
    bytes = 0;
}

int Staging::ensure()
{
    if (ready)
        return SDSLGPU_OK;
    for (int k = 0; k < kSlots; ++k)
    {
        SG_CUDA(cudaMalloc(reinterpret_cast<void **>(&in[k]), kChunk * kInBytesPerQuery));
        SG_CUDA(cudaMalloc(reinterpret_cast<void **>(&out[k]), kChunk * kOutBytesPerQuery));
        SG_CUDA(cudaStreamCreateWithFlags(&stream[k], cudaStreamNonBlocking));
    }
    ready = true;
    return SDSLGPU_OK;
}

void Staging::destroy()
{
    for (int k = 0; k < kSlots; ++k)
    {
        if (in[k])
            cudaFree(in[k]);
        if (out[k])
            cudaFree(out[k]);
        if (stream[k])
            cudaStreamDestroy(stream[k]);
        in[k] = out[k] = nullptr;
        stream[k] = nullptr;
    }
    ready = false;
}

int StagingIv::ensure()
{
    if (ready)
        return SDSLGPU_OK;
    for (int k = 0; k < kSlots; ++k)
    {
        SG_CUDA(cudaMalloc(reinterpret_cast<void **>(&pin[k]), (kChunk + 2) * 8));
        SG_CUDA(cudaMalloc(reinterpret_cast<void **>(&pout[k]), (kChunk + 2) * 8));
        SG_CUDA(cudaMalloc(reinterpret_cast<void **>(&uin[k]), kChunk * 8));
        SG_CUDA(cudaMalloc(reinterpret_cast<void **>(&uout[k]), kChunk * 8));
        SG_CUDA(cudaStreamCreateWithFlags(&stream[k], cudaStreamNonBlocking));
    }
    ready = true;
    return SDSLGPU_OK;
}

void StagingIv::destroy()
{
    for (int k = 0; k < kSlots; ++k)
    {
        for (uint64_t ** p : {&pin[k], &pout[k], &uin[k], &uout[k]})
        {
            if (*p)
                cudaFree(*p);
            *p = nullptr;
        }
        if (stream[k])
            cudaStreamDestroy(stream[k]);
        stream[k] = nullptr;
    }
    ready = false;
}

// ------------------------------------------------------------------------------------------------
// int_vector<w> wire format: field k occupies bits [k*w, (k+1)*w) of the word array, LSB first
// (int_vector.hpp get_int / bits::read_int, bits.hpp:777-790; bits::write_int :737-760)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) iv_unpack_kernel(uint64_t const * __restrict__ words, uint64_t nwords, uint32_t width, uint64_t n, uint64_t * __restrict__ out)
{
    uint64_t const stride = (uint64_t)gridDim.x * blockDim.x;
    uint64_t const mask = width >= 64 ? ~0ull : (1ull << width) - 1ull;
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride)
    {
        uint64_t const pos = k * width, w = pos >> 6;
        uint32_t const off = (uint32_t)(pos & 63);
        uint64_t v = words[w] >> off;
        if (off + width > 64 && w + 1 < nwords)
            v |= words[w + 1] << (64 - off);
        st_stream_u64(out + k, v & mask);
    }
}

// one thread per OUTPUT word: the (at most 64 / w + 2) fields overlapping it, each value truncated to w bits
// (SDSLGPU_NPOS becomes the all-ones field)
__global__ void __launch_bounds__(kThreads) iv_pack_kernel(uint64_t const * __restrict__ vals, uint64_t n, uint32_t width, uint64_t * __restrict__ words, uint64_t nwords)
{
    uint64_t const stride = (uint64_t)gridDim.x * blockDim.x;
    uint64_t const mask = width >= 64 ? ~0ull : (1ull << width) - 1ull;
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < nwords; j += stride)
    {
        uint64_t const lo_bit = j * 64, hi_bit = lo_bit + 64;
        uint64_t acc = 0;
        for (uint64_t k = lo_bit / width; k < n && k * width < hi_bit; ++k)
        {
            uint64_t const v = ld_stream_u64(vals + k) & mask, pos = k * width;
            acc |= pos >= lo_bit ? v << (pos - lo_bit) : v >> (lo_bit - pos);
        }
        words[j] = acc;
    }
}

static uint64_t iv_words(uint64_t n, uint32_t width)
{
    return (n * width + 63) >> 6;
}

// launch(in_u64_on_device, n, out_u64_on_device, stream) over queries / results that travel as int_vector<w> fields.
// Host arrays are streamed chunk by chunk (packed over PCIe, unpacked / packed on the device); device arrays are
// unpacked into stream-ordered scratch, answered and packed back, asynchronously on the caller's stream.
template <class Launch>
static int run_batch_iv(sdslgpu_handle const * hc, uint64_t const * in_words, uint32_t in_width, uint64_t n, uint64_t * out_words, uint32_t out_width,
                        cudaStream_t user, Launch launch)
{
    sdslgpu_handle * h = const_cast<sdslgpu_handle *>(hc);
    if (n == 0)
        return SDSLGPU_OK;
    if (!in_words || !out_words || in_width == 0 || in_width > 64 || out_width == 0 || out_width > 64)
    {
        set_error("int_vector batch: null pointer or width outside 1..64");
        return SDSLGPU_EINVAL;
    }
    DeviceGuard g(h->device);
    if (!g.ok)
    {
        set_error("cannot select CUDA device %d", h->device);
        return SDSLGPU_ECUDA;
    }
    PtrSpace isp, osp;
    SG_TRY(classify(in_words, h->device, &isp));
    SG_TRY(classify(out_words, h->device, &osp));
    if (isp == PtrSpace::Device && osp == PtrSpace::Device)
    {
        uint64_t * tmp = nullptr;
        SG_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&tmp), 2 * n * 8, user));
        iv_unpack_kernel<<<grid_for(n), kThreads, 0, user>>>(in_words, iv_words(n, in_width), in_width, n, tmp);
        int st = launch(tmp, n, tmp + n, user);
        if (st == SDSLGPU_OK)
        {
            uint64_t const ow = iv_words(n, out_width);
            iv_pack_kernel<<<grid_for(ow), kThreads, 0, user>>>(tmp + n, n, out_width, out_words, ow);
            if (cudaGetLastError() != cudaSuccess)
                st = SDSLGPU_ECUDA;
        }
        cudaFreeAsync(tmp, user);
        return st;
    }
    if (isp != osp)
    {
        set_error("int_vector batch: queries and results must both be host arrays or both device arrays");
        return SDSLGPU_EINVAL;
    }
    std::lock_guard<std::mutex> lock(h->staging.mu);
    SG_TRY(h->staging_iv.ensure());
    StagingIv & st = h->staging_iv;
    SG_CUDA(cudaStreamSynchronize(user));
    uint64_t const chunk = StagingIv::kChunk, nchunks = (n + chunk - 1) / chunk;
    uint64_t const in_total = iv_words(n, in_width), out_total = iv_words(n, out_width);
    int status = SDSLGPU_OK;
    for (uint64_t c = 0; c < nchunks && status == SDSLGPU_OK; ++c)
    {
        int const slot = (int)(c % StagingIv::kSlots);
        cudaStream_t s = st.stream[slot];
        uint64_t const lo = c * chunk, cnt = (n - lo < chunk) ? n - lo : chunk;
        // chunk = 2^23 queries: lo * width is a multiple of 64, so every chunk starts on a word boundary on both sides
        uint64_t const iw0 = lo * in_width / 64, ow0 = lo * out_width / 64;
        uint64_t iw = iv_words(cnt, in_width), ow = iv_words(cnt, out_width);
        if (iw0 + iw > in_total)
            iw = in_total - iw0;
        if (ow0 + ow > out_total)
            ow = out_total - ow0;
        cudaError_t e = cudaMemcpyAsync(st.pin[slot], in_words + iw0, iw * 8, cudaMemcpyHostToDevice, s);
        if (e != cudaSuccess)
        {
            status = cuda_fail(e, "H2D packed chunk", __FILE__, __LINE__);
            break;
        }
        iv_unpack_kernel<<<grid_for(cnt), kThreads, 0, s>>>(st.pin[slot], iw, in_width, cnt, st.uin[slot]);
        status = launch(st.uin[slot], cnt, st.uout[slot], s);
        if (status != SDSLGPU_OK)
            break;
        iv_pack_kernel<<<grid_for(ow), kThreads, 0, s>>>(st.uout[slot], cnt, out_width, st.pout[slot], ow);
        e = cudaGetLastError();
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(out_words + ow0, st.pout[slot], ow * 8, cudaMemcpyDeviceToHost, s);
        if (e != cudaSuccess)
            status = cuda_fail(e, "D2H packed chunk", __FILE__, __LINE__);
    }
    for (int k = 0; k < StagingIv::kSlots; ++k)
    {
        cudaError_t e = cudaStreamSynchronize(st.stream[k]);
        if (e != cudaSuccess && status == SDSLGPU_OK)
            status = cuda_fail(e, "staging sync", __FILE__, __LINE__);
    }
    return status;
}

int classify(void const * p, int device, PtrSpace * space)
{
    cudaPointerAttributes a;
    cudaError_t e = cudaPointerGetAttributes(&a, p);
    if (e != cudaSuccess)
    {
        cudaGetLastError();
        *space = PtrSpace::Host; // very old behaviour for unregistered memory
        return SDSLGPU_OK;
    }
    if (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged)
    {
        if (a.type == cudaMemoryTypeDevice && a.device != device)
        {
            set_error("device pointer %p lives on device %d but the handle is on device %d", p, a.device, device);
            return SDSLGPU_EINVAL;
        }
        *space = PtrSpace::Device;
    }
    else
        *space = PtrSpace::Host;
    return SDSLGPU_OK;
}

// ------------------------------------------------------------------------------------------------
// generic batch driver: `launch(in_ptrs_on_device, n, out_ptrs_on_device, stream)` for up to 2 input
// and 2 output columns.  If every pointer is a device pointer the launch is asynchronous on the
// caller's stream; otherwise host columns are streamed through the handle's staging slots.
// ------------------------------------------------------------------------------------------------
struct Column
{
    void const * in = nullptr; // user pointer (host or device)
    void * out = nullptr;
    uint32_t elem = 8; // bytes per element
};

template <class Launch>
static int run_batch(sdslgpu_handle const * hc, Column const * ins, int nin, Column const * outs, int nout, uint64_t n, cudaStream_t user, Launch launch)
{
    sdslgpu_handle * h = const_cast<sdslgpu_handle *>(hc);
    if (n == 0)
        return SDSLGPU_OK;
    DeviceGuard g(h->device);
    if (!g.ok)
    {
        set_error("cannot select CUDA device %d", h->device);
        return SDSLGPU_ECUDA;
    }
    bool any_host = false;
    PtrSpace in_sp[2] = {PtrSpace::Device, PtrSpace::Device}, out_sp[2] = {PtrSpace::Device, PtrSpace::Device};
    for (int k = 0; k < nin; ++k)
    {
        if (!ins[k].in)
        {
            set_error("null input pointer");
            return SDSLGPU_EINVAL;
        }
        SG_TRY(classify(ins[k].in, h->device, &in_sp[k]));
        any_host |= in_sp[k] == PtrSpace::Host;
    }
    for (int k = 0; k < nout; ++k)
    {
        if (!outs[k].out)
            continue; // optional output column
        SG_TRY(classify(outs[k].out, h->device, &out_sp[k]));
        any_host |= out_sp[k] == PtrSpace::Host;
    }
    if (!any_host)
    {
        void const * ip[2] = {nin > 0 ? ins[0].in : nullptr, nin > 1 ? ins[1].in : nullptr};
        void * op[2] = {nout > 0 ? outs[0].out : nullptr, nout > 1 ? outs[1].out : nullptr};
        return launch(ip, n, op, user);
    }

    // chunked pipeline over the staging slots
    std::lock_guard<std::mutex> lock(h->staging.mu);
    SG_TRY(h->staging.ensure());
    Staging & st = h->staging;
    // order after the caller's earlier work: the staging streams are non-blocking, so not even the legacy default
    // stream (user == NULL) is implicitly ordered before them
    SG_CUDA(cudaStreamSynchronize(user));
    uint64_t const chunk = Staging::kChunk;
    uint64_t nchunks = (n + chunk - 1) / chunk;
    int status = SDSLGPU_OK;
    for (uint64_t c = 0; c < nchunks && status == SDSLGPU_OK; ++c)
    {
        int slot = (int)(c % Staging::kSlots);
        cudaStream_t s = st.stream[slot];
        uint64_t lo = c * chunk, cnt = (n - lo < chunk) ? n - lo : chunk;
        void const * ip[2] = {nullptr, nullptr};
        void * op[2] = {nullptr, nullptr};
        // a slot is reused every kSlots chunks; same-stream ordering makes that safe
        for (int k = 0; k < nin; ++k)
        {
            uint8_t const * src = static_cast<uint8_t const *>(ins[k].in) + lo * ins[k].elem;
            if (in_sp[k] == PtrSpace::Host)
            {
                uint8_t * dst = st.in[slot] + (uint64_t)k * chunk * 8;
                cudaError_t e = cudaMemcpyAsync(dst, src, cnt * ins[k].elem, cudaMemcpyHostToDevice, s);
                if (e != cudaSuccess)
                    status = cuda_fail(e, "H2D chunk", __FILE__, __LINE__);
                ip[k] = dst;
            }
            else
                ip[k] = src;
        }
        for (int k = 0; k < nout; ++k)
        {
            if (!outs[k].out)
                continue;
            if (out_sp[k] == PtrSpace::Host)
                op[k] = st.out[slot] + (uint64_t)k * chunk * 8;
            else
                op[k] = static_cast<uint8_t *>(outs[k].out) + lo * outs[k].elem;
        }
        if (status == SDSLGPU_OK)
            status = launch(ip, cnt, op, s);
        for (int k = 0; k < nout && status == SDSLGPU_OK; ++k)
        {
            if (!outs[k].out || out_sp[k] != PtrSpace::Host)
                continue;
            uint8_t * dst = static_cast<uint8_t *>(outs[k].out) + lo * outs[k].elem;
            cudaError_t e = cudaMemcpyAsync(dst, op[k], cnt * outs[k].elem, cudaMemcpyDeviceToHost, s);
            if (e != cudaSuccess)
                status = cuda_fail(e, "D2H chunk", __FILE__, __LINE__);
        }
    }
    for (int k = 0; k < Staging::kSlots; ++k)
    {
        cudaError_t e = cudaStreamSynchronize(st.stream[k]);
        if (e != cudaSuccess && status == SDSLGPU_OK)
            status = cuda_fail(e, "staging sync", __FILE__, __LINE__);
    }
    return status;
}

static int check_handle(sdslgpu_handle const * h)
{
    if (!h)
    {
        set_error("null handle");
        return SDSLGPU_EINVAL;
    }
    return SDSLGPU_OK;
}

// Whole-buffer staging for calls with irregular (CSR) inputs such as pattern sets: host pointers are
// copied to temporary device buffers up front and host outputs copied back by finish(); device pointers
// pass through untouched (the call is then asynchronous on the caller's stream apart from the one
// 8-byte read-back locate needs to size its output).
struct StagedCall
{
    int device;
    cudaStream_t s;
    bool any_host = false;
    std::vector<void *> temps;
    struct Back
    {
        void * host;
        void * dev;
        uint64_t bytes;
    };
    std::vector<Back> backs;
    StagedCall(int dev, cudaStream_t st) : device(dev), s(st)
    {}
    ~StagedCall()
    {
        // stream-ordered frees: no device-wide synchronisation, the pool recycles the blocks for the next call
        for (void * p : temps)
            cudaFreeAsync(p, s);
    }
    int tmp(uint64_t bytes, void ** p)
    {
        cudaError_t e = cudaMallocAsync(p, bytes ? bytes : 8, s);
        if (e != cudaSuccess)
            return cuda_fail(e, "cudaMallocAsync (staging)", __FILE__, __LINE__);
        temps.push_back(*p);
        return SDSLGPU_OK;
    }
    template <class T>
    int in(T const * p, uint64_t bytes, T const ** dev)
    {
        PtrSpace sp = PtrSpace::Device;
        if (p == nullptr && bytes)
        {
            set_error("null input pointer");
            return SDSLGPU_EINVAL;
        }
        if (bytes)
            SG_TRY(classify(p, device, &sp));
        if (sp == PtrSpace::Device && bytes)
        {
            *dev = p;
            return SDSLGPU_OK;
        }
        void * d = nullptr;
        SG_TRY(tmp(bytes, &d));
        if (bytes)
        {
            any_host = true;
            SG_CUDA(cudaMemcpyAsync(d, p, bytes, cudaMemcpyHostToDevice, s));
        }
        *dev = static_cast<T const *>(d);
        return SDSLGPU_OK;
    }
    template <class T>
    int out(T * p, uint64_t bytes, T ** dev)
    {
        PtrSpace sp = PtrSpace::Host;
        SG_TRY(classify(p, device, &sp));
        if (sp == PtrSpace::Device)
        {
            *dev = p;
            return SDSLGPU_OK;
        }
        void * d = nullptr;
        SG_TRY(tmp(bytes, &d));
        any_host = true;
        backs.push_back(Back{p, d, bytes});
        *dev = static_cast<T *>(d);
        return SDSLGPU_OK;
    }
    template <class T>
    int scratch(uint64_t bytes, T ** dev)
    {
        void * d = nullptr;
        SG_TRY(tmp(bytes, &d));
        *dev = static_cast<T *>(d);
        return SDSLGPU_OK;
    }
    int peek_u64(uint64_t const * base, uint64_t index, uint64_t * value)
    {
        PtrSpace sp = PtrSpace::Host;
        SG_TRY(classify(base, device, &sp));
        if (sp == PtrSpace::Host)
        {
            *value = base[index];
            return SDSLGPU_OK;
        }
        SG_CUDA(cudaMemcpyAsync(value, base + index, 8, cudaMemcpyDeviceToHost, s));
        SG_CUDA(cudaStreamSynchronize(s));
        return SDSLGPU_OK;
    }
    int finish()
    {
        for (Back const & b : backs)
            if (b.bytes)
                SG_CUDA(cudaMemcpyAsync(b.host, b.dev, b.bytes, cudaMemcpyDeviceToHost, s));
        if (any_host || !temps.empty())
            SG_CUDA(cudaStreamSynchronize(s));
        backs.clear();
        return SDSLGPU_OK;
    }
};

static bool is_byte_wt(sdslgpu_handle const * h)
{
    return h->kind == SDSLGPU_KIND_WT_HUFF || h->kind == SDSLGPU_KIND_CSA_WT;
}

// keep freed scratch / staging blocks cached in the device's stream-ordered pool instead of returning them to the
// driver at every synchronisation (the default release threshold is 0: a 1 GB scratch would be re-mapped per call)
static void keep_stream_pool_cached(int device)
{
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess)
    {
        uint64_t keep = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    cudaGetLastError();
}

// validates the device (no CPU fallback) and allocates an empty handle of the given kind
static int new_handle(int kind, int device, uint32_t flags, sdslgpu_handle ** out)
{
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess)
    {
        cuda_fail(e, "cudaGetDeviceCount", __FILE__, __LINE__);
        return SDSLGPU_ECUDA;
    }
    if (device < 0 || device >= ndev)
    {
        set_error("device %d out of range (%d CUDA devices); there is no CPU fallback", device, ndev);
        return SDSLGPU_ECUDA;
    }
    keep_stream_pool_cached(device);
    sdslgpu_handle * h = new (std::nothrow) sdslgpu_handle;
    if (!h)
        return SDSLGPU_ENOMEM;
    h->kind = kind;
    h->device = device;
    h->flags = flags;
    *out = h;
    return SDSLGPU_OK;
}

} // namespace sdslgpu

using namespace sdslgpu;

// pattern codes: 0 / 1 everywhere; the two-bit patterns only on plain bit vectors
static int check_pattern(sdslgpu_handle const * h, int b, char const * who)
{
    if (b == 0 || b == 1)
        return SDSLGPU_OK;
    if (b >= SDSLGPU_PAT_10 && b <= SDSLGPU_PAT_11 && h->kind == SDSLGPU_KIND_BV)
        return SDSLGPU_OK;
    set_error("%s: pattern %d is not supported by this handle (0 / 1; SDSLGPU_PAT_10..11 on bit_vector handles)", who, b);
    return SDSLGPU_EINVAL;
}

// the indicator-vector image of a two-bit pattern, built on first use (the handle is logically const)
static int pattern_image(sdslgpu_handle const * ch, int b, BvImage const ** out)
{
    sdslgpu_handle * h = const_cast<sdslgpu_handle *>(ch);
    std::lock_guard<std::mutex> lock(h->pat_mu);
    int k = b - SDSLGPU_PAT_10;
    if (!h->pat_ready[k])
    {
        if (h->flags & SDSLGPU_F_NO_SELECT)
        {
            set_error("two-bit patterns need a handle created without SDSLGPU_F_NO_SELECT");
            return SDSLGPU_EINVAL;
        }
        DeviceGuard g(h->device);
        SG_TRY(bv_build_pattern(h->pool, h->bv, b, h->pat[k], nullptr));
        h->pat_ready[k] = true;
    }
    *out = &h->pat[k];
    return SDSLGPU_OK;
}

extern "C"
{

    const char * sdslgpu_version(void)
    {
        return "sdslgpu 0.1 (sm_100a; rank/select/wt/fm batched query engine)";
    }

    const char * sdslgpu_last_error(void)
    {
        return g_err;
    }

    int sdslgpu_device_count(int * count)
    {
        if (!count)
            return SDSLGPU_EINVAL;
        *count = 0;
        SG_CUDA(cudaGetDeviceCount(count));
        return SDSLGPU_OK;
    }

    int sdslgpu_bv_create(const uint64_t * words, uint64_t nbits, int device, uint32_t flags, sdslgpu_handle ** out)
    {
        if (!out || (!words && nbits))
        {
            set_error("sdslgpu_bv_create: null argument");
            return SDSLGPU_EINVAL;
        }
        *out = nullptr;
        int ndev = 0;
        SG_CUDA(cudaGetDeviceCount(&ndev));
        if (device < 0 || device >= ndev)
        {
            set_error("sdslgpu_bv_create: device %d out of range (%d devices); there is no CPU fallback", device, ndev);
            return SDSLGPU_ECUDA;
        }
        DeviceGuard g(device);
        if (!g.ok)
        {
            set_error("cannot select CUDA device %d", device);
            return SDSLGPU_ECUDA;
        }
        keep_stream_pool_cached(device);
        sdslgpu_handle * h = new (std::nothrow) sdslgpu_handle;
        if (!h)
            return SDSLGPU_ENOMEM;
        h->kind = SDSLGPU_KIND_BV;
        h->device = device;
        h->flags = flags;
        PtrSpace sp = PtrSpace::Host;
        int st = nbits ? classify(words, device, &sp) : SDSLGPU_OK;
        if (st == SDSLGPU_OK)
            st = bv_build(h->pool, h->bv, flags, words, sp == PtrSpace::Device, nbits, nullptr);
        if (st != SDSLGPU_OK)
        {
            h->pool.release_all();
            delete h;
            return st;
        }
        *out = h;
        return SDSLGPU_OK;
    }

    int sdslgpu_free(sdslgpu_handle * h)
    {
        if (!h)
            return SDSLGPU_OK;
        DeviceGuard g(h->device);
        cudaDeviceSynchronize();
        h->staging.destroy();
        h->staging_iv.destroy();
        h->pool.release_all();
        delete h;
        return SDSLGPU_OK;
    }

    int sdslgpu_kind(const sdslgpu_handle * h, int * kind)
    {
        SG_TRY(check_handle(h));
        *kind = h->kind;
        return SDSLGPU_OK;
    }

    int sdslgpu_size(const sdslgpu_handle * h, uint64_t * size)
    {
        SG_TRY(check_handle(h));
        switch (h->kind)
        {
        case SDSLGPU_KIND_BV:
            *size = h->bv.nbits;
            return SDSLGPU_OK;
        case SDSLGPU_KIND_WT_HUFF:
            *size = h->wt.size;
            return SDSLGPU_OK;
        case SDSLGPU_KIND_CSA_WT:
            *size = h->csa.n;
            return SDSLGPU_OK;
        case SDSLGPU_KIND_WT_INT:
            *size = h->wti.size;
            return SDSLGPU_OK;
        case SDSLGPU_KIND_RRR63:
            *size = h->rrr.size;
            return SDSLGPU_OK;
        case SDSLGPU_KIND_SD:
            *size = h->sd.size;
            return SDSLGPU_OK;
        }
        return SDSLGPU_ENOTSUP;
    }

    int sdslgpu_arg_count(const sdslgpu_handle * h, int b, uint64_t * count)
    {
        SG_TRY(check_handle(h));
        SG_TRY(check_pattern(h, b, "sdslgpu_arg_count"));
        switch (h->kind)
        {
        case SDSLGPU_KIND_BV:
            if (b >= 2)
            {
                BvImage const * img = nullptr;
                SG_TRY(pattern_image(h, b, &img));
                *count = img->ones;
                return SDSLGPU_OK;
            }
            *count = b ? h->bv.ones : h->bv.nbits - h->bv.ones;
            return SDSLGPU_OK;
        case SDSLGPU_KIND_RRR63:
            *count = b ? h->rrr.ones : h->rrr.size - h->rrr.ones;
            return SDSLGPU_OK;
        case SDSLGPU_KIND_SD:
            *count = b ? h->sd.m : h->sd.size - h->sd.m;
            return SDSLGPU_OK;
        }
        return SDSLGPU_ENOTSUP;
    }

    int sdslgpu_device_bytes(const sdslgpu_handle * h, uint64_t * bytes)
    {
        SG_TRY(check_handle(h));
        *bytes = h->pool.bytes;
        return SDSLGPU_OK;
    }

    int sdslgpu_set_batch_order(sdslgpu_handle * h, int order)
    {
        SG_TRY(check_handle(h));
        if (h->kind != SDSLGPU_KIND_BV && h->kind != SDSLGPU_KIND_SD && h->kind != SDSLGPU_KIND_RRR63)
        {
            set_error("sdslgpu_set_batch_order: only bit-vector handles (plain, sd, rrr) have a binned batch path");
            return SDSLGPU_ENOTSUP;
        }
        if (order != SDSLGPU_ORDER_AUTO && order != SDSLGPU_ORDER_DIRECT && order != SDSLGPU_ORDER_BINNED)
        {
            set_error("sdslgpu_set_batch_order: unknown order %d", order);
            return SDSLGPU_EINVAL;
        }
        std::lock_guard<std::mutex> lock(h->pat_mu);
        h->order = order;
        h->bv.order = order;
        for (int k = 0; k < 4; ++k)
            h->pat[k].order = order;
        return SDSLGPU_OK;
    }

    int sdslgpu_auto_is_binned(uint64_t index_bytes, uint64_t n, int select)
    {
        return bin_wanted(SDSLGPU_ORDER_AUTO, index_bytes, n, select ? kBinSelectDensity : kBinRankDensity) ? 1 : 0;
    }

    int sdslgpu_rank(const sdslgpu_handle * h, int b, const uint64_t * idx, uint64_t n, uint64_t * out, void * stream)
    {
        SG_TRY(check_handle(h));
        SG_TRY(check_pattern(h, b, "sdslgpu_rank"));
        if (n && !out)
        {
            set_error("sdslgpu_rank: null output");
            return SDSLGPU_EINVAL;
        }
        Column in{idx, nullptr, 8}, o{nullptr, out, 8};
        switch (h->kind)
        {
        case SDSLGPU_KIND_BV:
            if (b >= 2)
            { // rank_support_v<pattern, 2>: one-bit rank over the pattern's indicator vector
                BvImage const * img = nullptr;
                SG_TRY(pattern_image(h, b, &img));
                return run_batch(h, &in, 1, &o, 1, n, static_cast<cudaStream_t>(stream), [=](void const * const * ip, uint64_t cnt, void * const * op, cudaStream_t s) {
                    return bv_rank_device(*img, SDSLGPU_F_DEFAULT, 1, static_cast<uint64_t const *>(ip[0]), cnt, static_cast<uint64_t *>(op[0]), s);
                });
            }
            return run_batch(h, &in, 1, &o, 1, n, static_cast<cudaStream_t>(stream), [=](void const * const * ip, uint64_t cnt, void * const * op, cudaStream_t s) {
                return bv_rank_device(h->bv, h->flags, b, static_cast<uint64_t const *>(ip[0]), cnt, static_cast<uint64_t *>(op[0]), s);
            });
        case SDSLGPU_KIND_RRR63:
            return run_batch(h, &in, 1, &o, 1, n, static_cast<cudaStream_t>(stream), [=](void const * const * ip, uint64_t cnt, void * const * op, cudaStream_t s) {
                return rrr_rank_device(h, b, static_cast<uint64_t const *>(ip[0]), cnt, static_cast<uint64_t *>(op[0]), s);
            });
        case SDSLGPU_KIND_SD:
            return run_batch(h, &in, 1, &o, 1, n, static_cast<cudaStream_t>(stream), [=](void const * const * ip, uint64_t cnt, void * const * op, cudaStream_t s) {
                return sd_rank_device(h, b, static_cast<uint64_t const *>(ip[0]), cnt, static_cast<uint64_t *>(op[0]), s);
            });
        }
        set_error("sdslgpu_rank: unsupported handle kind %d", h->kind);
        return SDSLGPU_ENOTSUP;
    }

    int sdslgpu_select(const sdslgpu_handle * h, int b, const uint64_t * i, uint64_t n, uint64_t * out, void * stream)
    {
        SG_TRY(check_handle(h));
        SG_TRY(check_pattern(h, b, "sdslgpu_select"));
        if (n && !out)
        {
            set_error("sdslgpu_select: null output");
            return SDSLGPU_EINVAL;
        }
        Column in{i, nullptr, 8}, o{nullptr, out, 8};
        switch (h->kind)
        {
        case SDSLGPU_KIND_BV:
            if (b >= 2)
            {
                BvImage const * img = nullptr;
                SG_TRY(pattern_image(h, b, &img));
                return run_batch(h, &in, 1, &o, 1, n, static_cast<cudaStream_t>(stream), [=](void const * const * ip, uint64_t cnt, void * const * op, cudaStream_t s) {
                    return bv_select_device(*img, 1, static_cast<uint64_t const *>(ip[0]), cnt, static_cast<uint64_t *>(op[0]), s);
                });
            }
            return run_batch(h, &in, 1, &o, 1, n, static_cast<cudaStream_t>(stream), [=](void const * const * ip, uint64_t cnt, void * const * op, cudaStream_t s) {
                SG_TRY(bv_ensure_select_sectors(h, b, cnt)); // first large batch: one-gather select (bv_device.cuh)
                return bv_select_device(h->bv, b, static_cast<uint64_t const *>(ip[0]), cnt, static_cast<uint64_t *>(op[0]), s);
            });
        case SDSLGPU_KIND_RRR63:
            return run_batch(h, &in, 1, &o, 1, n, static_cast<cudaStream_t>(stream), [=](void const * const * ip, uint64_t cnt, void * const * op, cudaStream_t s) {
                return rrr_select_device(h, b, static_cast<uint64_t const *>(ip[0]), cnt, static_cast<uint64_t *>(op[0]), s);
            });
        case SDSLGPU_KIND_SD:
            return run_batch(h, &in, 1, &o, 1, n, static_cast<cudaStream_t>(stream), [=](void const * const * ip, uint64_t cnt, void * const * op, cudaStream_t s) {
                return sd_select_device(h, b, static_cast<uint64_t const *>(ip[0]), cnt, static_cast<uint64_t *>(op[0]), s);
            });
        }
        set_error("sdslgpu_select: unsupported handle kind %d", h->kind);
        return SDSLGPU_ENOTSUP;
    }

    // queries / results as int_vector<w> fields.  op: 0 = rank, 1 = select
    static int rank_select_iv(const sdslgpu_handle * h, int op, int b, const uint64_t * in_words, uint32_t in_width, uint64_t n, uint64_t * out_words,
                              uint32_t out_width, void * stream)
    {
        SG_TRY(check_handle(h));
        char const * who = op ? "sdslgpu_select_iv" : "sdslgpu_rank_iv";
        if (b != 0 && b != 1)
        {
            set_error("%s: b must be 0 or 1", who);
            return SDSLGPU_EINVAL;
        }
        cudaStream_t user = static_cast<cudaStream_t>(stream);
        switch (h->kind)
        {
        case SDSLGPU_KIND_BV:
            return run_batch_iv(h, in_words, in_width, n, out_words, out_width, user, [=](uint64_t const * q, uint64_t cnt, uint64_t * o, cudaStream_t s) {
                if (op)
                    SG_TRY(bv_ensure_select_sectors(h, b, cnt));
                return op ? bv_select_device(h->bv, b, q, cnt, o, s) : bv_rank_device(h->bv, h->flags, b, q, cnt, o, s);
            });
        case SDSLGPU_KIND_RRR63:
            return run_batch_iv(h, in_words, in_width, n, out_words, out_width, user, [=](uint64_t const * q, uint64_t cnt, uint64_t * o, cudaStream_t s) {
                return op ? rrr_select_device(h, b, q, cnt, o, s) : rrr_rank_device(h, b, q, cnt, o, s);
            });
        case SDSLGPU_KIND_SD:
            return run_batch_iv(h, in_words, in_width, n, out_words, out_width, user, [=](uint64_t const * q, uint64_t cnt, uint64_t * o, cudaStream_t s) {
                return op ? sd_select_device(h, b, q, cnt, o, s) : sd_rank_device(h, b, q, cnt, o, s);
            });
        }
        set_error("%s: unsupported handle kind %d", who, h->kind);
        return SDSLGPU_ENOTSUP;
    }

    int sdslgpu_rank_iv(const sdslgpu_handle * h, int b, const uint64_t * idx_words, uint32_t idx_width, uint64_t n, uint64_t * out_words, uint32_t out_width,
                        void * stream)
    {
        return rank_select_iv(h, 0, b, idx_words, idx_width, n, out_words, out_width, stream);
    }

    int sdslgpu_select_iv(const sdslgpu_handle * h, int b, const uint64_t * i_words, uint32_t i_width, uint64_t n, uint64_t * out_words, uint32_t out_width,
                          void * stream)
    {
        return rank_select_iv(h, 1, b, i_words, i_width, n, out_words, out_width, stream);
    }

    int sdslgpu_access(const sdslgpu_handle * h, const uint64_t * idx, uint64_t n, uint64_t * out, void * stream)
    {
        SG_TRY(check_handle(h));
        if (n && !out)
        {
            set_error("sdslgpu_access: null output");
            return SDSLGPU_EINVAL;
        }
        Column in{idx, nullptr, 8}, o{nullptr, out, 8};
        switch (h->kind)
        {
        case SDSLGPU_KIND_BV:
            return run_batch(h, &in, 1, &o, 1, n, static_cast<cudaStream_t>(stream), [=](void const * const * ip, uint64_t cnt, void * const * op, cudaStream_t s) {
                return bv_access_device(h->bv, static_cast<uint64_t const *>(ip[0]), cnt, static_cast<uint64_t *>(op[0]), s);
            });
        case SDSLGPU_KIND_RRR63:
            return run_batch(h, &in, 1, &o, 1, n, static_cast<cudaStream_t>(stream), [=](void const * const * ip, uint64_t cnt, void * const * op, cudaStream_t s) {
                return rrr_access_device(h, static_cast<uint64_t const *>(ip[0]), cnt, static_cast<uint64_t *>(op[0]), s);
            });
        case SDSLGPU_KIND_SD:
            return run_batch(h, &in, 1, &o, 1, n, static_cast<cudaStream_t>(stream), [=](void const * const * ip, uint64_t cnt, void * const * op, cudaStream_t s) {
                return sd_access_device(h, static_cast<uint64_t const *>(ip[0]), cnt, static_cast<uint64_t *>(op[0]), s);
            });
        }
        set_error("sdslgpu_access: unsupported handle kind %d", h->kind);
        return SDSLGPU_ENOTSUP;
    }

    int sdslgpu_bv_serialize(const sdslgpu_handle * h, int what, void * buf, uint64_t cap, uint64_t * nbytes)
    {
        SG_TRY(check_handle(h));
        if (h->kind != SDSLGPU_KIND_BV || !nbytes)
            return SDSLGPU_EINVAL;
        if (!(h->flags & SDSLGPU_F_SDSL_LAYOUT))
        {
            set_error("sdslgpu_bv_serialize needs a handle created with SDSLGPU_F_SDSL_LAYOUT");
            return SDSLGPU_ENOTSUP;
        }
        BvImage const & v = h->bv;
        uint64_t header, words;
        uint64_t const * src;
        if (what == 0)
        {
            header = (1ull << 56) | v.nbits; // int_vector.hpp:904-916
            words = v.nwords;
            src = v.words;
        }
        else if (what == 1 || what == 2)
        {
            words = v.table_words;
            header = (64ull << 56) | (words * 64);
            src = v.rank_table[what == 1 ? 1 : 0];
        }
        else
            return SDSLGPU_EINVAL;
        *nbytes = 8 + words * 8;
        if (!buf)
            return SDSLGPU_OK;
        if (cap < *nbytes)
        {
            set_error("sdslgpu_bv_serialize: buffer too small");
            return SDSLGPU_EINVAL;
        }
        DeviceGuard g(h->device);
        std::memcpy(buf, &header, 8);
        if (words)
            SG_CUDA(cudaMemcpy(static_cast<uint8_t *>(buf) + 8, src, words * 8, cudaMemcpyDeviceToHost));
        return SDSLGPU_OK;
    }

    // -------------------------------------------------------------------------------- compressed bit vectors
    static int create_compressed(int kind, const uint64_t * words, uint64_t nbits, int device, uint32_t flags, sdslgpu_handle ** out)
    {
        if (!out || (!words && nbits))
        {
            set_error("create: null argument");
            return SDSLGPU_EINVAL;
        }
        *out = nullptr;
        sdslgpu_handle * h = nullptr;
        SG_TRY(new_handle(kind, device, flags, &h));
        DeviceGuard g(device);
        PtrSpace sp = PtrSpace::Host;
        int st = nbits ? classify(words, device, &sp) : SDSLGPU_OK;
        if (st == SDSLGPU_OK)
            st = (kind == SDSLGPU_KIND_RRR63) ? rrr_build(h, words, sp == PtrSpace::Device, nbits, nullptr)
                                              : sd_build(h, words, sp == PtrSpace::Device, nbits, nullptr);
        if (st != SDSLGPU_OK)
        {
            h->pool.release_all();
            delete h;
            return st;
        }
        *out = h;
        return SDSLGPU_OK;
    }

    int sdslgpu_rrr63_create(const uint64_t * words, uint64_t nbits, int device, uint32_t flags, sdslgpu_handle ** out)
    {
        return create_compressed(SDSLGPU_KIND_RRR63, words, nbits, device, flags, out);
    }

    int sdslgpu_sd_create(const uint64_t * words, uint64_t nbits, int device, uint32_t flags, sdslgpu_handle ** out)
    {
        return create_compressed(SDSLGPU_KIND_SD, words, nbits, device, flags, out);
    }

    int sdslgpu_load_sdsl_ex(const void * blob, uint64_t nbytes, int kind, int device, uint32_t flags, uint32_t sa_dens, uint32_t isa_dens,
                             uint64_t * consumed, sdslgpu_handle ** out)
    {
        if (!out || !blob)
        {
            set_error("sdslgpu_load_sdsl: null argument");
            return SDSLGPU_EINVAL;
        }
        *out = nullptr;
        if (kind < SDSLGPU_KIND_BV || kind > SDSLGPU_KIND_CSA_WT)
        {
            set_error("sdslgpu_load_sdsl: unknown kind %d", kind);
            return SDSLGPU_EINVAL;
        }
        sdslgpu_handle * h = nullptr;
        SG_TRY(new_handle(kind, device, flags, &h));
        DeviceGuard g(device);
        int st = load_sdsl_blob(h, static_cast<uint8_t const *>(blob), nbytes, sa_dens, isa_dens, consumed, nullptr);
        if (st != SDSLGPU_OK)
        {
            h->pool.release_all();
            delete h;
            return st;
        }
        *out = h;
        return SDSLGPU_OK;
    }

    int sdslgpu_load_sdsl(const void * blob, uint64_t nbytes, int kind, int device, uint32_t flags, uint32_t param, sdslgpu_handle ** out)
    {
        return sdslgpu_load_sdsl_ex(blob, nbytes, kind, device, flags, param, 0, nullptr, out);
    }

    int sdslgpu_serialize(const sdslgpu_handle * h, int what, void * buf, uint64_t cap, uint64_t * nbytes)
    {
        SG_TRY(check_handle(h));
        if (!nbytes)
            return SDSLGPU_EINVAL;
        if (h->kind == SDSLGPU_KIND_BV && what >= 0 && what <= 2 && (h->flags & SDSLGPU_F_SDSL_LAYOUT))
            return sdslgpu_bv_serialize(h, what, buf, cap, nbytes); // resident words / tables
        DeviceGuard g(h->device);
        sdslgpu_handle * hm = const_cast<sdslgpu_handle *>(h);
        std::lock_guard<std::mutex> lock(hm->ser_mu);
        std::vector<uint8_t> blob;
        bool const fetch = hm->ser_what == what && buf;
        if (!fetch && hm->ser_what != -1)
        { // a size query whose blob was never fetched: dropped by the next call instead of living as long as the handle
            std::vector<uint8_t>().swap(hm->ser_blob);
            hm->ser_what = -1;
        }
        if (fetch)
        { // the size query before this call already built it
            blob.swap(hm->ser_blob);
            hm->ser_what = -1;
        }
        else if (h->kind == SDSLGPU_KIND_BV && ((what >= 0 && what <= 2) || what == 5 || what == 6))
            SG_TRY(egress_bv_part(h->bv, what, blob)); // unpacked from the sector blocks, table rebuilt on the device
        else if (h->kind == SDSLGPU_KIND_BV && (what == 3 || what == 4))
            SG_TRY(egress_select_mcl(h->bv, what == 3 ? 1 : 0, blob));
        else if (h->kind == SDSLGPU_KIND_RRR63 && what == 0)
            SG_TRY(rrr_serialize(h, blob));
        else if (h->kind == SDSLGPU_KIND_SD && what == 0)
            SG_TRY(sd_serialize_low_high(h, blob));
        else if (h->kind == SDSLGPU_KIND_SD && what == 1)
            SG_TRY(egress_sd(h, blob));
        else if (h->kind == SDSLGPU_KIND_WT_HUFF && (what == 0 || what == 1))
            SG_TRY(egress_wt_huff(h, blob, what == 1));
        else if (h->kind == SDSLGPU_KIND_WT_INT && what == 0)
            SG_TRY(egress_wt_int(h, blob));
        else if (h->kind == SDSLGPU_KIND_CSA_WT && (what == 0 || what == 1))
            SG_TRY(egress_csa(h, blob, what == 1));
        else
        {
            set_error("sdslgpu_serialize: unsupported (kind %d, what %d)", h->kind, what);
            return SDSLGPU_ENOTSUP;
        }
        *nbytes = blob.size();
        if (!buf)
        {
            hm->ser_blob.swap(blob);
            hm->ser_what = what;
            return SDSLGPU_OK;
        }
        if (cap < blob.size())
        {
            set_error("sdslgpu_serialize: buffer too small");
            return SDSLGPU_EINVAL;
        }
        std::memcpy(buf, blob.data(), blob.size());
        return SDSLGPU_OK;
    }

    // -------------------------------------------------------------------------------- wavelet trees
    int sdslgpu_wt_huff_create(const uint8_t * text, uint64_t n, int device, uint32_t flags, sdslgpu_handle ** out)
    {
        if (!out || (!text && n))
        {
            set_error("sdslgpu_wt_huff_create: null argument");
            return SDSLGPU_EINVAL;
        }
        *out = nullptr;
        sdslgpu_handle * h = nullptr;
        SG_TRY(new_handle(SDSLGPU_KIND_WT_HUFF, device, flags, &h));
        DeviceGuard g(device);
        int st = wt_huff_build_from_text(h, text, n, nullptr);
        if (st != SDSLGPU_OK)
        {
            h->pool.release_all();
            delete h;
            return st;
        }
        *out = h;
        return SDSLGPU_OK;
    }

    int sdslgpu_wt_int_create(const uint64_t * seq, uint64_t n, int device, uint32_t flags, sdslgpu_handle ** out)
    {
        if (!out || (!seq && n))
        {
            set_error("sdslgpu_wt_int_create: null argument");
            return SDSLGPU_EINVAL;
        }
        *out = nullptr;
        sdslgpu_handle * h = nullptr;
        SG_TRY(new_handle(SDSLGPU_KIND_WT_INT, device, flags, &h));
        DeviceGuard g(device);
        int st = wt_int_build(h, seq, n, nullptr);
        if (st != SDSLGPU_OK)
        {
            h->pool.release_all();
            delete h;
            return st;
        }
        *out = h;
        return SDSLGPU_OK;
    }

    int sdslgpu_wt_sigma(const sdslgpu_handle * h, uint64_t * sigma)
    {
        SG_TRY(check_handle(h));
        if (!sigma || !(is_byte_wt(h) || h->kind == SDSLGPU_KIND_WT_INT))
            return SDSLGPU_EINVAL;
        *sigma = h->kind == SDSLGPU_KIND_WT_INT ? h->wti.sigma : h->wt.sigma;
        return SDSLGPU_OK;
    }

    int sdslgpu_wt_rank(const sdslgpu_handle * h, const uint64_t * i, const void * c, uint64_t n, uint64_t * out, void * stream)
    {
        SG_TRY(check_handle(h));
        if (n && (!out || !c))
        {
            set_error("sdslgpu_wt_rank: null argument");
            return SDSLGPU_EINVAL;
        }
        if (is_byte_wt(h))
        {
            Column in[2] = {{i, nullptr, 8}, {c, nullptr, 1}};
            Column o{nullptr, out, 8};
            return run_batch(h, in, 2, &o, 1, n, static_cast<cudaStream_t>(stream), [=](void const * const * ip, uint64_t cnt, void * const * op, cudaStream_t s) {
                return wt_rank_device(h, static_cast<uint64_t const *>(ip[0]), static_cast<uint8_t const *>(ip[1]), cnt, static_cast<uint64_t *>(op[0]), s);
            });
        }
        if (h->kind == SDSLGPU_KIND_WT_INT)
        {
            Column in[2] = {{i, nullptr, 8}, {c, nullptr, 8}};
            Column o{nullptr, out, 8};
            return run_batch(h, in, 2, &o, 1, n, static_cast<cudaStream_t>(stream), [=](void const * const * ip, uint64_t cnt, void * const * op, cudaStream_t s) {
                return wt_int_rank_device(h, static_cast<uint64_t const *>(ip[0]), static_cast<uint64_t const *>(ip[1]), cnt, static_cast<uint64_t *>(op[0]), s);
            });
        }
        set_error("sdslgpu_wt_rank: unsupported handle kind %d", h->kind);
        return SDSLGPU_ENOTSUP;
    }

    int sdslgpu_wt_select(const sdslgpu_handle * h, const uint64_t * i, const void * c, uint64_t n, uint64_t * out, void * stream)
    {
        SG_TRY(check_handle(h));
        if (n && (!out || !c))
        {
            set_error("sdslgpu_wt_select: null argument");
            return SDSLGPU_EINVAL;
        }
        if (is_byte_wt(h))
        {
            Column in[2] = {{i, nullptr, 8}, {c, nullptr, 1}};
            Column o{nullptr, out, 8};
            return run_batch(h, in, 2, &o, 1, n, static_cast<cudaStream_t>(stream), [=](void const * const * ip, uint64_t cnt, void * const * op, cudaStream_t s) {
                return wt_select_device(h, static_cast<uint64_t const *>(ip[0]), static_cast<uint8_t const *>(ip[1]), cnt, static_cast<uint64_t *>(op[0]), s);
            });
        }
        if (h->kind == SDSLGPU_KIND_WT_INT)
        {
            Column in[2] = {{i, nullptr, 8}, {c, nullptr, 8}};
            Column o{nullptr, out, 8};
            return run_batch(h, in, 2, &o, 1, n, static_cast<cudaStream_t>(stream), [=](void const * const * ip, uint64_t cnt, void * const * op, cudaStream_t s) {
                return wt_int_select_device(h, static_cast<uint64_t const *>(ip[0]), static_cast<uint64_t const *>(ip[1]), cnt, static_cast<uint64_t *>(op[0]), s);
            });
        }
        set_error("sdslgpu_wt_select: unsupported handle kind %d", h->kind);
        return SDSLGPU_ENOTSUP;
    }

    int sdslgpu_wt_access(const sdslgpu_handle * h, const uint64_t * i, uint64_t n, uint64_t * sym_out, uint64_t * rank_out, void * stream)
    {
        SG_TRY(check_handle(h));
        if (n && !sym_out)
        {
            set_error("sdslgpu_wt_access: null output");
            return SDSLGPU_EINVAL;
        }
        if (is_byte_wt(h))
        {
            Column in{i, nullptr, 8};
            Column o[2] = {{nullptr, sym_out, 8}, {nullptr, rank_out, 8}};
            return run_batch(h, &in, 1, o, 2, n, static_cast<cudaStream_t>(stream), [=](void const * const * ip, uint64_t cnt, void * const * op, cudaStream_t s) {
                return wt_access_device(h, static_cast<uint64_t const *>(ip[0]), cnt, static_cast<uint64_t *>(op[0]), static_cast<uint64_t *>(op[1]), s);
            });
        }
        if (h->kind == SDSLGPU_KIND_WT_INT)
        {
            Column in{i, nullptr, 8};
            Column o[2] = {{nullptr, sym_out, 8}, {nullptr, rank_out, 8}};
            return run_batch(h, &in, 1, o, 2, n, static_cast<cudaStream_t>(stream), [=](void const * const * ip, uint64_t cnt, void * const * op, cudaStream_t s) {
                return wt_int_access_device(h, static_cast<uint64_t const *>(ip[0]), cnt, static_cast<uint64_t *>(op[0]), static_cast<uint64_t *>(op[1]), s);
            });
        }
        set_error("sdslgpu_wt_access: unsupported handle kind %d", h->kind);
        return SDSLGPU_ENOTSUP;
    }

    // -------------------------------------------------------------------------------- FM-index
    int sdslgpu_csa_create_ex(const uint8_t * text, uint64_t n, int device, uint32_t flags, uint32_t sa_dens, uint32_t isa_dens, sdslgpu_handle ** out)
    {
        if (!out || (!text && n))
        {
            set_error("sdslgpu_csa_create: null argument");
            return SDSLGPU_EINVAL;
        }
        *out = nullptr;
        sdslgpu_handle * h = nullptr;
        SG_TRY(new_handle(SDSLGPU_KIND_CSA_WT, device, flags, &h));
        DeviceGuard g(device);
        int st = csa_build_from_text(h, text, n, sa_dens, isa_dens, nullptr);
        if (st != SDSLGPU_OK)
        {
            h->pool.release_all();
            delete h;
            return st;
        }
        *out = h;
        return SDSLGPU_OK;
    }

    int sdslgpu_csa_create(const uint8_t * text, uint64_t n, int device, uint32_t flags, sdslgpu_handle ** out)
    {
        return sdslgpu_csa_create_ex(text, n, device, flags, 0, 0, out);
    }

    int sdslgpu_csa_alphabet(const sdslgpu_handle * h, uint64_t * C, uint8_t * char2comp, uint8_t * comp2char, uint32_t * sigma)
    {
        SG_TRY(check_handle(h));
        if (h->kind != SDSLGPU_KIND_CSA_WT)
        {
            set_error("sdslgpu_csa_alphabet: handle is not a CSA");
            return SDSLGPU_ENOTSUP;
        }
        FmTables const & t = h->csa.host_tab;
        if (C)
            std::memcpy(C, t.C, sizeof(t.C));
        if (char2comp)
            std::memcpy(char2comp, t.char2comp, 256);
        if (comp2char)
            std::memcpy(comp2char, t.comp2char, 256);
        if (sigma)
            *sigma = t.sigma;
        return SDSLGPU_OK;
    }

    int sdslgpu_fm_count(const sdslgpu_handle * h, const uint8_t * pats, const uint64_t * pat_off, uint64_t n, uint64_t * cnt_out, uint64_t * l_out, void * stream)
    {
        SG_TRY(check_handle(h));
        if (h->kind != SDSLGPU_KIND_CSA_WT)
        {
            set_error("sdslgpu_fm_count: handle is not a CSA");
            return SDSLGPU_ENOTSUP;
        }
        if (n == 0)
            return SDSLGPU_OK;
        if (!pat_off || !cnt_out)
        {
            set_error("sdslgpu_fm_count: null argument");
            return SDSLGPU_EINVAL;
        }
        DeviceGuard g(h->device);
        cudaStream_t s = static_cast<cudaStream_t>(stream);
        StagedCall sc(h->device, s);
        uint64_t const * d_off = nullptr;
        uint8_t const * d_pats = nullptr;
        uint64_t *d_cnt = nullptr, *d_l = nullptr;
        SG_TRY(sc.in(pat_off, (n + 1) * 8, &d_off));
        uint64_t total_bytes = 0;
        SG_TRY(sc.peek_u64(pat_off, n, &total_bytes));
        SG_TRY(sc.in(pats, total_bytes, &d_pats));
        SG_TRY(sc.out(cnt_out, n * 8, &d_cnt));
        if (l_out)
            SG_TRY(sc.out(l_out, n * 8, &d_l));
        SG_TRY(fm_count_device(h, d_pats, d_off, n, d_cnt, d_l, s));
        return sc.finish();
    }

    int sdslgpu_fm_sa(const sdslgpu_handle * h, const uint64_t * i, uint64_t n, uint64_t * out, void * stream)
    {
        SG_TRY(check_handle(h));
        if (h->kind != SDSLGPU_KIND_CSA_WT)
        {
            set_error("sdslgpu_fm_sa: handle is not a CSA");
            return SDSLGPU_ENOTSUP;
        }
        if (n && (!i || !out))
        {
            set_error("sdslgpu_fm_sa: null argument");
            return SDSLGPU_EINVAL;
        }
        Column in{i, nullptr, 8}, o{nullptr, out, 8};
        return run_batch(h, &in, 1, &o, 1, n, static_cast<cudaStream_t>(stream), [=](void const * const * ip, uint64_t cnt, void * const * op, cudaStream_t s) {
            return fm_sa_device(h, static_cast<uint64_t const *>(ip[0]), cnt, static_cast<uint64_t *>(op[0]), s);
        });
    }

    int sdslgpu_fm_extract(const sdslgpu_handle * h, const uint64_t * begin, const uint64_t * end, uint64_t n, const uint64_t * out_off, uint8_t * out, void * stream)
    {
        SG_TRY(check_handle(h));
        if (h->kind != SDSLGPU_KIND_CSA_WT)
        {
            set_error("sdslgpu_fm_extract: handle is not a CSA");
            return SDSLGPU_ENOTSUP;
        }
        if (n == 0)
            return SDSLGPU_OK;
        if (!begin || !end || !out_off || !out)
        {
            set_error("sdslgpu_fm_extract: null argument");
            return SDSLGPU_EINVAL;
        }
        DeviceGuard g(h->device);
        cudaStream_t s = static_cast<cudaStream_t>(stream);
        StagedCall sc(h->device, s);
        uint64_t const *d_b = nullptr, *d_e = nullptr, *d_off = nullptr;
        uint8_t * d_out = nullptr;
        uint64_t total = 0;
        SG_TRY(sc.in(begin, n * 8, &d_b));
        SG_TRY(sc.in(end, n * 8, &d_e));
        SG_TRY(sc.in(out_off, (n + 1) * 8, &d_off));
        SG_TRY(sc.peek_u64(out_off, n, &total));
        SG_TRY(sc.out(out, total, &d_out));
        SG_TRY(fm_extract_device(h, d_b, d_e, d_off, n, d_out, s));
        return sc.finish();
    }

    int sdslgpu_fm_locate(const sdslgpu_handle * h,
                          const uint8_t * pats,
                          const uint64_t * pat_off,
                          uint64_t n,
                          uint64_t * occ_off_out,
                          uint64_t * occ_out,
                          uint64_t occ_cap,
                          uint64_t * total_out,
                          void * stream)
    {
        SG_TRY(check_handle(h));
        if (h->kind != SDSLGPU_KIND_CSA_WT)
        {
            set_error("sdslgpu_fm_locate: handle is not a CSA");
            return SDSLGPU_ENOTSUP;
        }
        if (!occ_off_out || !total_out || (n && !pat_off))
        {
            set_error("sdslgpu_fm_locate: null argument");
            return SDSLGPU_EINVAL;
        }
        DeviceGuard g(h->device);
        cudaStream_t s = static_cast<cudaStream_t>(stream);
        StagedCall sc(h->device, s);
        uint64_t const * d_off = nullptr;
        uint8_t const * d_pats = nullptr;
        uint64_t *d_cnt = nullptr, *d_l = nullptr, *d_occ_off = nullptr, *d_tmp = nullptr, *d_occ = nullptr;
        uint64_t total_bytes = 0, total = 0;
        if (n)
        {
            SG_TRY(sc.in(pat_off, (n + 1) * 8, &d_off));
            SG_TRY(sc.peek_u64(pat_off, n, &total_bytes));
            SG_TRY(sc.in(pats, total_bytes, &d_pats));
        }
        SG_TRY(sc.scratch((n + 1) * 8, &d_cnt));
        SG_TRY(sc.scratch((n + 1) * 8, &d_l));
        SG_TRY(sc.scratch(fm_scan_tmp_words(n) * 8, &d_tmp));
        SG_TRY(sc.out(occ_off_out, (n + 1) * 8, &d_occ_off));
        SG_TRY(fm_count_device(h, d_pats, d_off, n, d_cnt, d_l, s));
        SG_TRY(fm_scan_counts_device(d_cnt, n, d_occ_off, d_tmp, s));
        SG_CUDA(cudaMemcpyAsync(&total, d_occ_off + n, 8, cudaMemcpyDeviceToHost, s));
        SG_CUDA(cudaStreamSynchronize(s));
        *total_out = total;
        if (occ_out && total)
        {
            if (occ_cap < total)
            {
                sc.finish();
                set_error("sdslgpu_fm_locate: occ_out holds %llu entries but %llu occurrences were found", (unsigned long long)occ_cap,
                          (unsigned long long)total);
                return SDSLGPU_EINVAL;
            }
            SG_TRY(sc.out(occ_out, total * 8, &d_occ));
            SG_TRY(fm_locate_fill_device(h, d_l, d_occ_off, n, total, d_occ, s));
        }
        return sc.finish();
    }

} // extern "C"
