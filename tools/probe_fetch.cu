// tools/probe_fetch.cu — how many DRAM bytes does ONE random 32-byte gather cost on B200, per load flavour?
// (ncu on probe_gather showed 4 DRAM sectors = 128 B per 8..32-byte gather with the default ld.global.nc path
// and cudaLimitMaxL2FetchGranularity ignored.)  Run under ncu with dram__bytes_read.sum; each variant is a
// separate kernel name.  Also measures random 8-byte SCATTER stores.  Not part of the product library.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <cuda/barrier>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__global__ void fill(uint64_t* p, uint64_t n, uint64_t seed) {
  uint64_t s = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += s) {
    uint64_t x = (i + seed) * 0x9E3779B97F4A7C15ull; x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 32; x *= 0x94D049BB133111EBull; x ^= x >> 29;
    p[i] = x;
  }
}

#define GATHER_KERNEL(NAME, LOADSTMT)                                                                                   \
  __global__ void __launch_bounds__(256) NAME(const uint8_t* __restrict__ base, uint64_t units, const uint64_t* __restrict__ idx, uint64_t n, uint64_t* __restrict__ out) { \
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;                                                                 \
    for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += stride) {                            \
      const uint8_t* p = base + (idx[q] % units) * 32;                                                                  \
      uint32_t r0 = 0, r1 = 0, r2 = 0, r3 = 0;                                                                          \
      LOADSTMT;                                                                                                         \
      out[q] = r0 ^ r1 ^ r2 ^ r3;                                                                                       \
    }                                                                                                                   \
  }

GATHER_KERNEL(g_nc_noalloc_v4, asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "l"(p)))
GATHER_KERNEL(g_plain_v4, asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "l"(p)))
GATHER_KERNEL(g_cg_v4, asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "l"(p)))
GATHER_KERNEL(g_cv_v4, asm volatile("ld.global.cv.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "l"(p)))
GATHER_KERNEL(g_lu_v4, asm volatile("ld.global.lu.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "l"(p)))
GATHER_KERNEL(g_cs_v4, asm volatile("ld.global.cs.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "l"(p)))
GATHER_KERNEL(g_L2_64B_v4, asm volatile("ld.global.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "l"(p)))
GATHER_KERNEL(g_evict_first_v4, { uint64_t a; uint64_t b; uint64_t c; uint64_t d; asm volatile("ld.global.L1::no_allocate.L2::evict_first.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p)); r0 = (uint32_t)(a ^ b ^ c ^ d); })
GATHER_KERNEL(g_relaxed_gpu_b64, { uint64_t v; asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p)); r0 = (uint32_t)v; r1 = (uint32_t)(v >> 32); })
GATHER_KERNEL(g_atom_or0_b32, { asm volatile("atom.global.or.b32 %0, [%1], 0;" : "=r"(r0) : "l"(p)); })

// LDGSTS: cp.async 16 B per thread into shared memory
__global__ void __launch_bounds__(256) g_cp_async16(const uint8_t* __restrict__ base, uint64_t units, const uint64_t* __restrict__ idx, uint64_t n, uint64_t* __restrict__ out) {
  __shared__ __align__(16) uint32_t sm[256 * 4];
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += stride) {
    const uint8_t* p = base + (idx[q] % units) * 32;
    uint32_t dst = (uint32_t)__cvta_generic_to_shared(sm + threadIdx.x * 4);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(p));
    asm volatile("cp.async.commit_group;");
    asm volatile("cp.async.wait_group 0;");
    out[q] = sm[threadIdx.x * 4] ^ sm[threadIdx.x * 4 + 3];
  }
}

// TMA-style bulk copy (UBLKCP): every thread fetches its own 32-byte block into shared memory
__global__ void __launch_bounds__(256) g_bulk32(const uint8_t* __restrict__ base, uint64_t units, const uint64_t* __restrict__ idx, uint64_t n, uint64_t* __restrict__ out) {
  __shared__ __align__(128) uint32_t sm[256 * 8];
  __shared__ __align__(8) uint64_t bar;
  uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&bar);
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a)); }
  __syncthreads();
  uint32_t phase = 0;
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  uint64_t nround = (n + stride - 1) / stride;
  for (uint64_t rd = 0; rd < nround; ++rd) {
    uint64_t q = rd * stride + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (threadIdx.x == 0) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(256u * 32u)); }
    __syncthreads();
    const uint8_t* p = base + ((q < n ? idx[q] : 0) % units) * 32;
    uint32_t dst = (uint32_t)__cvta_generic_to_shared(sm + threadIdx.x * 8);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 32, [%2];" ::"r"(dst), "l"(p), "r"(bar_a) : "memory");
    uint32_t done = 0;
    while (!done) {
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar_a), "r"(phase) : "memory");
    }
    phase ^= 1;
    if (q < n) out[q] = sm[threadIdx.x * 8] ^ sm[threadIdx.x * 8 + 7];
    __syncthreads();
  }
}

// random 8-byte scatter stores into a region of `units` x 8 bytes
__global__ void __launch_bounds__(256) s_scatter8(uint64_t* __restrict__ dst, uint64_t units, const uint64_t* __restrict__ idx, uint64_t n) {
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += stride) dst[idx[q] % units] = q;
}
// a permutation scatter: every 8-byte slot written exactly once (idx = bijective hash of q), the out[qid] pattern
__global__ void __launch_bounds__(256) s_permute8(uint64_t* __restrict__ dst, uint64_t log2n, uint64_t n) {
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x, mask = (1ull << log2n) - 1;
  for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += stride) {
    uint64_t x = q; // bijection on log2n bits: odd multiply + xorshift rounds
    x = (x * 0x9E3779B97F4A7C15ull) & mask; x ^= x >> (log2n / 2); x = (x * 0xBF58476D1CE4E5B9ull) & mask; x ^= x >> (log2n / 2 + 1);
    x = (x * 0x94D049BB133111EBull) & mask;
    dst[x] = q;
  }
}

template <class F>
void timeit(const char* name, uint64_t n, F f) {
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  float best = 1e30f;
  for (int it = 0; it < 4; ++it) {
    CK(cudaEventRecord(a)); f(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b)); if (it >= 1 && ms < best) best = ms;
  }
  CK(cudaGetLastError());
  printf("%-22s %8.3f ms  %7.2f G/s\n", name, best, n / (best * 1e-3) / 1e9);
}

int main(int argc, char** argv) {
  uint64_t n = argc > 1 ? (uint64_t)atof(argv[1]) : (1ull << 26);
  uint64_t bytes = 1234567936ull;  // ~1.15 GiB
  uint8_t* base; uint64_t *idx, *out;
  CK(cudaMalloc(&base, bytes + 256)); CK(cudaMalloc(&idx, n * 8)); CK(cudaMalloc(&out, n * 8));
  fill<<<148 * 8, 256>>>((uint64_t*)base, bytes / 8, 1); fill<<<148 * 8, 256>>>(idx, n, 3);
  CK(cudaDeviceSynchronize());
  uint64_t units = bytes / 32; int grid = 148 * 8;
#define RUN(K) timeit(#K, n, [&] { K<<<grid, 256>>>(base, units, idx, n, out); })
  RUN(g_nc_noalloc_v4); RUN(g_plain_v4); RUN(g_cg_v4); RUN(g_cv_v4); RUN(g_lu_v4); RUN(g_cs_v4); RUN(g_L2_64B_v4); RUN(g_evict_first_v4);
  RUN(g_relaxed_gpu_b64); RUN(g_atom_or0_b32); RUN(g_cp_async16); RUN(g_bulk32);
  timeit("s_scatter8 (1.15 GiB)", n, [&] { s_scatter8<<<grid, 256>>>((uint64_t*)base, bytes / 8, idx, n); });
  timeit("s_permute8 (n x 8 B)", n, [&] { s_permute8<<<grid, 256>>>(out, 26, n); });
  timeit("s_scatter8 (64 MiB)", n, [&] { s_scatter8<<<grid, 256>>>((uint64_t*)base, (64ull << 20) / 8, idx, n); });
  return 0;
}
