// internal.h — host-side handle layout and helpers shared by the C-ABI translation units.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/sdslgpu.h"
#include "common.cuh"

namespace sdslgpu
{

void set_error(char const * fmt, ...);
int cuda_fail(cudaError_t e, char const * what, char const * file, int line);

#define SG_CUDA(expr)                                                                                                  \
    do                                                                                                                 \
    {                                                                                                                  \
        cudaError_t e__ = (expr);                                                                                      \
        if (e__ != cudaSuccess)                                                                                        \
            return ::sdslgpu::cuda_fail(e__, #expr, __FILE__, __LINE__);                                               \
    } while (0)

#define SG_TRY(expr)                                                                                                   \
    do                                                                                                                 \
    {                                                                                                                  \
        int s__ = (expr);                                                                                              \
        if (s__ != SDSLGPU_OK)                                                                                         \
            return s__;                                                                                                \
    } while (0)

// grid sizing: B200 has 148 SMs; query kernels are latency-bound gathers, so run a grid-stride loop
// over 148 x kCtasPerSm resident CTAs of 256 threads (full occupancy at <= 32 registers/thread).
static constexpr int kSmCount = 148;
static constexpr int kThreads = 256;

struct DeviceGuard
{
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev)
    {
        if (cudaGetDevice(&prev) == cudaSuccess && cudaSetDevice(dev) == cudaSuccess)
            ok = true;
    }
    ~DeviceGuard()
    {
        if (prev >= 0)
            cudaSetDevice(prev);
    }
};

// Owns device allocations of one handle and keeps the byte total.
struct DevicePool
{
    std::vector<void *> ptrs;
    std::vector<uint64_t> sizes_;
    uint64_t bytes = 0;
    int alloc(void ** p, uint64_t n);
    template <class T>
    int alloc_t(T ** p, uint64_t count)
    {
        return alloc(reinterpret_cast<void **>(p), count * sizeof(T));
    }
    void release(void * p); // free one allocation early (build temporaries)
    void release_all();
};

// Pinned-free staging path for HOST query buffers: device chunk buffers + two streams, so that the
// H2D copy of chunk k+1, the kernel of chunk k and the D2H copy of chunk k-1 overlap (PCIe is full duplex).
struct Staging
{
    static constexpr int kSlots = 3;
    static constexpr uint64_t kChunk = 1ull << 22; // queries per chunk
    std::mutex mu;
    bool ready = false;
    uint8_t * in[kSlots] = {nullptr, nullptr, nullptr};  // kChunk * in_bytes_max
    uint8_t * out[kSlots] = {nullptr, nullptr, nullptr}; // kChunk * out_bytes_max
    cudaStream_t stream[kSlots] = {nullptr, nullptr, nullptr};
    static constexpr uint64_t kInBytesPerQuery = 16; // largest input record (i:u64 + c:u64)
    static constexpr uint64_t kOutBytesPerQuery = 16;
    int ensure();
    void destroy();
};

enum class PtrSpace
{
    Host,
    Device
};
// classifies a user pointer; device pointers must belong to `device`
int classify(void const * p, int device, PtrSpace * space);

} // namespace sdslgpu

// ------------------------------------------------------------------------------------------------
// The opaque handle.  `kind` selects which of the per-kind images is populated.
// ------------------------------------------------------------------------------------------------
struct sdslgpu_bv_image
{
    uint64_t nbits = 0;
    uint64_t nblocks = 0;               // nbits/224 + 1 (one past the end so rank(size) stays in range)
    sdslgpu::bvblock * blocks = nullptr; // sector-interleaved payload + counts
    uint64_t * top = nullptr;            // absolute 1-count per superblock of 2^24 blocks
    uint64_t ntop = 0;
    uint64_t ones = 0;
    // select samples, per pattern b: samp[b][j] = block holding the (j*S+1)-th b-bit; one sentinel
    uint32_t * samp[2] = {nullptr, nullptr};
    uint64_t nsamp[2] = {0, 0};
    uint32_t log_s[2] = {6, 6};
    // optional SDSL layout (SDSLGPU_F_SDSL_LAYOUT): raw words (+1 pad word) and m_basic_block tables
    uint64_t * words = nullptr;
    uint64_t nwords = 0;
    uint64_t * rank_table[2] = {nullptr, nullptr}; // [b]
    uint64_t table_words = 0;
};

struct sdslgpu_handle
{
    int kind = 0;
    int device = 0;
    uint32_t flags = 0;
    sdslgpu::DevicePool pool;
    sdslgpu::Staging staging;
    sdslgpu_bv_image bv;
};

namespace sdslgpu
{
// bv.cu
int bv_build(sdslgpu_handle * h, uint64_t const * words_host_or_dev, bool words_on_device, uint64_t nbits, cudaStream_t s);
int bv_rank_device(sdslgpu_handle const * h, int b, uint64_t const * idx, uint64_t n, uint64_t * out, cudaStream_t s);
int bv_select_device(sdslgpu_handle const * h, int b, uint64_t const * idx, uint64_t n, uint64_t * out, cudaStream_t s);
int bv_access_device(sdslgpu_handle const * h, uint64_t const * idx, uint64_t n, uint64_t * out, cudaStream_t s);
int bv_build_sdsl_rank_table(sdslgpu_handle * h, int b, cudaStream_t s);
} // namespace sdslgpu
