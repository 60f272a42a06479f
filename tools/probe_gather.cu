// tools/probe_gather.cu — microbenchmark behind DESIGN.md §3: what does a uniformly random gather cost on
// B200 as a function of the bytes fetched per query?  Decides the rank block size (32-byte sector blocks vs
// 64 / 128-byte lines vs the reference's two-gather layout).  Not part of the product library.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/probe_gather tools/probe_gather.cu
//   tools/bin/probe_gather [log2_bytes=30.3] [queries=1e8]
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t ld32B(const void* p) {
  uint32_t r[8];
  asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
   : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]) : "l"(p));
  return r[0]^r[1]^r[2]^r[3]^r[4]^r[5]^r[6]^r[7];
}
__device__ __forceinline__ uint32_t ld16B(const void* p) {
  uint32_t r[4];
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]) : "l"(p));
  return r[0]^r[1]^r[2]^r[3];
}
__device__ __forceinline__ uint32_t ld8B(const void* p) {
  uint32_t r[2];
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r[0]),"=r"(r[1]) : "l"(p));
  return r[0]^r[1];
}

// MODE: bytes gathered per query at ONE random aligned location (8,16,32,64,128), or
//       24 = the reference's layout: 16 B from a table at idx/512*16 and 8 B from words at idx/64*8
template <int MODE, int ILP>
__global__ void __launch_bounds__(256) gather(const uint8_t* __restrict__ base, uint64_t span_units, const uint8_t* __restrict__ base2,
                                              const uint64_t* __restrict__ idx, uint64_t n, uint64_t* __restrict__ out) {
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x * ILP;
  for (uint64_t b = (uint64_t)blockIdx.x * blockDim.x * ILP + threadIdx.x; b < n; b += stride) {
    uint64_t i[ILP]; uint32_t acc[ILP];
#pragma unroll
    for (int u = 0; u < ILP; ++u) { uint64_t q = b + (uint64_t)u * blockDim.x; i[u] = q < n ? idx[q] : 0; }
#pragma unroll
    for (int u = 0; u < ILP; ++u) {
      if (MODE == 24) {
        uint64_t bit = i[u] % (span_units * 64);                 // span_units = number of 64-bit words
        acc[u] = ld16B(base2 + (bit >> 9) * 16) ^ ld8B(base + (bit >> 6) * 8);
      } else {
        const uint8_t* p = base + (i[u] % span_units) * (MODE < 32 ? 32 : MODE);   // <32: one partial sector
        if (MODE == 8) acc[u] = ld8B(p);
        if (MODE == 16) acc[u] = ld16B(p);
        if (MODE == 32) acc[u] = ld32B(p);
        if (MODE == 64) acc[u] = ld32B(p) ^ ld32B(p + 32);
        if (MODE == 128) acc[u] = ld32B(p) ^ ld32B(p + 32) ^ ld32B(p + 64) ^ ld32B(p + 96);
      }
    }
#pragma unroll
    for (int u = 0; u < ILP; ++u) { uint64_t q = b + (uint64_t)u * blockDim.x; if (q < n) out[q] = acc[u]; }
  }
}

__global__ void fill(uint64_t* p, uint64_t n, uint64_t seed) {
  uint64_t s = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += s) {
    uint64_t x = (i + seed) * 0x9E3779B97F4A7C15ull; x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 32; x *= 0x94D049BB133111EBull; x ^= x >> 29;
    p[i] = x;
  }
}

template <int MODE, int ILP>
void run(const char* name, const uint8_t* base, uint64_t bytes, const uint8_t* base2, const uint64_t* idx, uint64_t n, uint64_t* out, int ctas_per_sm) {
  uint64_t units = (MODE == 24) ? bytes / 8 : bytes / (MODE < 32 ? 32 : MODE);
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  int grid = 148 * ctas_per_sm;
  float best = 1e30f;
  for (int it = 0; it < 6; ++it) {
    CK(cudaEventRecord(a));
    gather<MODE, ILP><<<grid, 256>>>(base, units, base2, idx, n, out);
    CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    if (it >= 2 && ms < best) best = ms;
  }
  CK(cudaGetLastError());
  double qps = n / (best * 1e-3);
  double gathered = (MODE == 24 ? 24.0 : MODE) * qps / 1e9, sectors = (MODE == 24 ? 64.0 : (MODE < 32 ? 32 : MODE)) * qps / 1e9;
  printf("%-28s ilp=%d cta/sm=%d  %8.3f ms  %7.2f Gq/s  gathered %7.1f GB/s  sector-bytes %7.1f GB/s  (+16 B/q stream: %6.1f GB/s)\n",
         name, ILP, ctas_per_sm, best, qps / 1e9, gathered, sectors, 16 * qps / 1e9);
}

int main(int argc, char** argv) {
  double lg = argc > 1 ? atof(argv[1]) : 30.2;
  if (argc > 3) { cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, atoi(argv[3])); printf("pre-set granularity %s: %s\n", argv[3], cudaGetErrorString(e)); }
  uint64_t n = argc > 2 ? (uint64_t)atof(argv[2]) : 100000000ull;
  uint64_t bytes = ((uint64_t)exp2(lg)) & ~255ull;
  uint8_t *base, *base2; uint64_t *idx, *out;
  CK(cudaMalloc(&base, bytes + 256)); CK(cudaMalloc(&base2, bytes / 4 + 256));
  CK(cudaMalloc(&idx, n * 8)); CK(cudaMalloc(&out, n * 8));
  fill<<<148 * 8, 256>>>((uint64_t*)base, bytes / 8, 1); fill<<<148 * 8, 256>>>((uint64_t*)base2, bytes / 32, 2);
  fill<<<148 * 8, 256>>>(idx, n, 3);
  CK(cudaDeviceSynchronize());
  printf("footprint %.3f GiB, %llu queries (8 B in + 8 B out streamed per query)\n", bytes / 1073741824.0, (unsigned long long)n);
  run<8, 1>("1 x 8 B", base, bytes, base2, idx, n, out, 8);
  run<16, 1>("1 x 16 B", base, bytes, base2, idx, n, out, 8);
  run<32, 1>("1 x 32 B (sector block)", base, bytes, base2, idx, n, out, 8);
  run<32, 2>("1 x 32 B (sector block)", base, bytes, base2, idx, n, out, 8);
  run<32, 4>("1 x 32 B (sector block)", base, bytes, base2, idx, n, out, 8);
  run<32, 2>("1 x 32 B (sector block)", base, bytes, base2, idx, n, out, 16);
  run<64, 1>("1 x 64 B", base, bytes, base2, idx, n, out, 8);
  run<64, 2>("1 x 64 B", base, bytes, base2, idx, n, out, 8);
  run<128, 1>("1 x 128 B (line)", base, bytes, base2, idx, n, out, 8);
  run<128, 2>("1 x 128 B (line)", base, bytes, base2, idx, n, out, 8);
  run<24, 1>("16 B table + 8 B word (SDSL)", base, bytes, base2, idx, n, out, 8);
  run<24, 2>("16 B table + 8 B word (SDSL)", base, bytes, base2, idx, n, out, 8);
  run<24, 4>("16 B table + 8 B word (SDSL)", base, bytes, base2, idx, n, out, 8);
  // L2 fetch granularity: ncu shows the default path reading FOUR sectors from DRAM per 32-byte gather
  for (int gran : {32, 64, 128}) {
    cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran);
    size_t got = 0; cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
    printf("cudaLimitMaxL2FetchGranularity <- %d : %s, reads back %zu\n", gran, cudaGetErrorString(e), got);
    run<32, 2>("1 x 32 B (sector block)", base, bytes, base2, idx, n, out, 8);
    run<64, 2>("1 x 64 B", base, bytes, base2, idx, n, out, 8);
    run<128, 2>("1 x 128 B (line)", base, bytes, base2, idx, n, out, 8);
    run<128, 4>("1 x 128 B (line)", base, bytes, base2, idx, n, out, 8);
    run<24, 2>("16 B table + 8 B word (SDSL)", base, bytes, base2, idx, n, out, 8);
  }
  cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, 32);
  printf("footprint sweep at granularity 32\n");
  for (double g : {0.125, 0.25, 0.5, 1.0, 2.0, 4.0, 8.0}) {
    uint64_t fb = (uint64_t)(g * 1073741824.0);
    if (fb > bytes) break;
    printf("footprint %.3f GiB: ", g);
    run<32, 2>("1 x 32 B", base, fb, base2, idx, n, out, 8);
  }
  // L2-resident footprint for contrast
  uint64_t small = 64ull << 20;
  printf("footprint 64 MiB (L2 resident)\n");
  run<32, 2>("1 x 32 B (sector block)", base, small, base2, idx, n, out, 8);
  run<128, 2>("1 x 128 B (line)", base, small, base2, idx, n, out, 8);
  return 0;
}
