// tools/probe_line.cu — cost of ONE random gather from a 2 GiB region on B200 as a function of how much of the
// 128-byte line it touches: 32 B (one sector), 2 B + 32 B in the two sectors of a 64-byte half line, 64 B, 128 B —
// with and without the L2::64B fill hint, and with / without the 16 B/query index and result streams.
// Decides the record size of the FM-index occurrence structure.  Not part of the product library.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint64_t mix(uint64_t x) {
  x *= 0x9E3779B97F4A7C15ull; x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 32; x *= 0x94D049BB133111EBull; x ^= x >> 29;
  return x;
}
__global__ void fill(uint64_t* p, uint64_t n, uint64_t seed) {
  uint64_t s = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += s) p[i] = mix(i + seed);
}

__device__ __forceinline__ uint32_t ld256(const uint8_t* p) {
  uint32_t a, b, c, d, e, f, g, h;
  asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d), "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(p));
  return a ^ b ^ c ^ d ^ e ^ f ^ g ^ h;
}
__device__ __forceinline__ uint32_t ld256h(const uint8_t* p) {
  uint32_t a, b, c, d, e, f, g, h;
  asm volatile("ld.global.nc.L1::no_allocate.L2::64B.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d), "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(p));
  return a ^ b ^ c ^ d ^ e ^ f ^ g ^ h;
}
__device__ __forceinline__ uint32_t ld16h(const uint8_t* p) {
  uint16_t v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::64B.u16 %0, [%1];" : "=h"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ uint32_t ld16(const uint8_t* p) {
  uint16_t v;
  asm volatile("ld.global.nc.L1::no_allocate.u16 %0, [%1];" : "=h"(v) : "l"(p));
  return v;
}

__device__ __forceinline__ uint32_t ld256a(const uint8_t* p) {  // allocating in L1
  uint32_t a, b, c, d, e, f, g, h;
  asm volatile("ld.global.nc.L2::64B.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d), "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(p));
  return a ^ b ^ c ^ d ^ e ^ f ^ g ^ h;
}
__device__ __forceinline__ uint32_t ld16a(const uint8_t* p) {
  uint16_t v;
  asm volatile("ld.global.nc.L2::64B.u16 %0, [%1];" : "=h"(v) : "l"(p));
  return v;
}

// MODE: 0 = 32 B hint, 1 = 2 B + 32 B (two sectors of a half line) hint, 2 = 64 B hint (2 x 256-bit), 3 = 128 B (4 x 256-bit, no hint)
//       4 = 32 B no hint, 5 = 2 B + 32 B no hint, 6 = 64 B no hint, 7 = 2 B + 16 B of the SAME sector (hint)
//       8 = the same 32 B twice (two instructions), 9 = 32 B then a DEPENDENT 2 B from the other sector, 10 = dependent 2 B same sector,
//       11 = mode 1 with L1-allocating loads, 12 = 2 B only, 13 = dependent 2 B other sector, L1-allocating
template <int MODE, bool STREAMS>
__global__ void __launch_bounds__(256) gather(const uint8_t* __restrict__ base, uint64_t lines, const uint64_t* __restrict__ idx, uint64_t n, uint64_t* __restrict__ out, uint32_t zero = 0) {
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  uint32_t acc = 0;
  for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += stride) {
    uint64_t r = STREAMS ? idx[q] : mix(q);
    const uint8_t* p = base + (r % lines) * 128 + ((r >> 40) & 1) * 64;  // a random 64-byte half line
    uint32_t v = (r >> 50) & 15, x;
    if (MODE == 0) x = ld256h(p);
    if (MODE == 1) x = ld16h(p + 2 * v) + ld256h(p + 32);
    if (MODE == 2) x = ld256h(p) ^ ld256h(p + 32);
    if (MODE == 3) { p = base + (r % lines) * 128; x = ld256(p) ^ ld256(p + 32) ^ ld256(p + 64) ^ ld256(p + 96); }
    if (MODE == 4) x = ld256(p);
    if (MODE == 5) x = ld16(p + 2 * v) + ld256(p + 32);
    if (MODE == 6) x = ld256(p) ^ ld256(p + 32);
    if (MODE == 7) x = ld16h(p + 2 * (v & 7)) + ld256h(p);
    if (MODE == 8) x = ld256h(p) + ld256h(p);
    if (MODE == 9) { x = ld256h(p + 32); x += ld16h(p + 2 * v + (x & zero)); }
    if (MODE == 10) { x = ld256h(p); x += ld16h(p + 2 * (v & 7) + (x & zero)); }
    if (MODE == 11) x = ld16a(p + 2 * v) + ld256a(p + 32);
    if (MODE == 12) x = ld16h(p + 2 * v);
    if (MODE == 13) { x = ld256a(p + 32); x += ld16a(p + 2 * v + (x & zero)); }
    if (STREAMS) out[q] = x; else acc ^= x;
  }
  if (!STREAMS && acc == 0x12345678u) out[0] = acc;
}

// two DEPENDENT half-line gathers per query (level 1 -> level 2 of a 16-ary structure)
template <int MODE>
__global__ void __launch_bounds__(256) chain2(const uint8_t* __restrict__ base, uint64_t lines, const uint64_t* __restrict__ idx, uint64_t n, uint64_t* __restrict__ out) {
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += stride) {
    uint64_t r = idx[q];
    const uint8_t* p = base + (r % lines) * 128 + ((r >> 40) & 1) * 64;
    uint32_t v = (r >> 50) & 15;
    uint32_t x = MODE == 1 ? ld16h(p + 2 * v) + ld256h(p + 32) : ld256h(p);
    uint64_t r2 = mix(r + x);
    p = base + (r2 % lines) * 128 + ((r2 >> 40) & 1) * 64;
    x ^= MODE == 1 ? ld16h(p + 2 * v) + ld256h(p + 32) : ld256h(p);
    out[q] = x;
  }
}

// TLB test: all 256 threads of a CTA gather (32 B each) inside ONE region of 2^region_log2 bytes per iteration; the
// regions themselves are picked at random over the whole 2 GiB, so DRAM / L2 see the same random sector traffic.
// warp_local: the region is chosen per warp instead of per CTA.
__global__ void __launch_bounds__(256) gather_local(const uint8_t* __restrict__ base, uint64_t bytes, uint32_t region_log2, int warp_local, uint64_t n, uint64_t* __restrict__ out) {
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  uint64_t regions = bytes >> region_log2, per = (1ull << region_log2) / 32;
  uint64_t it = 0;
  for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += stride, ++it) {
    uint64_t who = warp_local ? (q >> 5) : (it * gridDim.x + blockIdx.x);
    uint64_t reg = mix(who * 2 + 12345) % regions;
    const uint8_t* p = base + (reg << region_log2) + (mix(q) % per) * 32;
    out[q] = ld256h(p);
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
  printf("%-44s %8.3f ms  %7.2f G gathers/s\n", name, best, n / (best * 1e-3) / 1e9);
}

int main(int argc, char** argv) {
  uint64_t n = argc > 1 ? (uint64_t)atof(argv[1]) : (1ull << 26);
  uint64_t bytes = 2ull << 30;
  uint8_t* base; uint64_t *idx, *out;
  CK(cudaMalloc(&base, bytes + 256)); CK(cudaMalloc(&idx, n * 8)); CK(cudaMalloc(&out, n * 8));
  fill<<<148 * 8, 256>>>((uint64_t*)base, bytes / 8, 1); fill<<<148 * 8, 256>>>(idx, n, 3);
  CK(cudaDeviceSynchronize());
  uint64_t lines = bytes / 128; int grid = 148 * 8;
#define RUN(M, S, LABEL) timeit(LABEL, n, [&] { gather<M, S><<<grid, 256>>>(base, lines, idx, n, out); })
  RUN(0, true, "32 B, L2::64B, streams");
  RUN(7, true, "2 B + 32 B same sector, L2::64B, streams");
  RUN(1, true, "2 B + 32 B two sectors, L2::64B, streams");
  RUN(2, true, "64 B, L2::64B, streams");
  RUN(4, true, "32 B, default fill, streams");
  RUN(5, true, "2 B + 32 B two sectors, default, streams");
  RUN(6, true, "64 B, default fill, streams");
  RUN(3, true, "128 B, default fill, streams");
  RUN(8, true, "32 B twice (same address), L2::64B, streams");
  RUN(9, true, "32 B then dependent 2 B other sector");
  RUN(10, true, "32 B then dependent 2 B same sector");
  RUN(11, true, "2 B + 32 B two sectors, L1-allocating");
  RUN(12, true, "2 B only");
  RUN(13, true, "32 B then dependent 2 B other sector, L1-alloc");
  RUN(0, false, "32 B, L2::64B, no streams");
  RUN(1, false, "2 B + 32 B two sectors, L2::64B, no streams");
  RUN(2, false, "64 B, L2::64B, no streams");
  RUN(3, false, "128 B, default fill, no streams");
  for (int wl = 0; wl < 2; ++wl)
    for (uint32_t rl : {16u, 21u, 23u, 25u, 27u, 29u, 31u}) {
      char label[96]; snprintf(label, sizeof label, "32 B, all lanes of a %s in one 2^%u B region", wl ? "warp" : "CTA", rl);
      timeit(label, n, [&] { gather_local<<<grid, 256>>>(base, bytes, rl, wl, n, out); });
    }
  // footprint sweep: uniformly random 32 B gathers over the first F bytes
  for (uint64_t f : {64ull << 20, 128ull << 20, 256ull << 20, 512ull << 20, 1024ull << 20, 2048ull << 20}) {
    char label[96]; snprintf(label, sizeof label, "32 B, uniform over %llu MiB", (unsigned long long)(f >> 20));
    uint32_t fl = 63 - __builtin_clzll(f);
    timeit(label, n, [&] { gather_local<<<grid, 256>>>(base, f, fl, 0, n, out); });
  }
  timeit("chain of 2 dependent 32 B gathers", n, [&] { chain2<0><<<grid, 256>>>(base, lines, idx, n, out); });
  timeit("chain of 2 dependent (2 B + 32 B) gathers", n, [&] { chain2<1><<<grid, 256>>>(base, lines, idx, n, out); });
  return 0;
}
