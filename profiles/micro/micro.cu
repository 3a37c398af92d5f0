// micro.cu -- B200 micro-measurements that size the MSDeformAttn kernels (not part of the product).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o micro micro.cu && ./micro
// 1. gather rate of 64-byte chunks (one bf16 head of one pixel) through L1 (LDG.128) and from
//    shared memory (LDS.128), random chunk positions inside a per-CTA window;
// 2. the same for 128-byte contiguous "row pairs" (two x-adjacent pixels of a head-major layout);
// 3. scatter rate of red.global.add.v4.f32 / .v4.bf16x2 into a 176 MB buffer and an L2-sized one;
// 4. shared-memory atomicAdd(float) rate, random addresses.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ unsigned rng(unsigned& s) { s = s * 1664525u + 1013904223u; return s >> 8; }

// mode 0: 4 lanes per 64-byte chunk (8 chunks per warp request); chunk index random in window
// mode 1: 8 lanes per 128-byte pair (4 pairs per warp request); pair start random at 64-byte granularity
template <int MODE, bool SMEM>
__global__ void __launch_bounds__(256) gather_kernel(const uint4* __restrict__ src, int window_chunks, int iters, float* sink) {
  extern __shared__ uint4 sm[];
  const uint4* base = src + (size_t)blockIdx.x * window_chunks * 4;  // 64-byte chunks = 4 uint4
  if (SMEM) {
    for (int i = threadIdx.x; i < window_chunks * 4; i += blockDim.x) sm[i] = base[i];
    __syncthreads();
    base = sm;
  }
  // mode 2: 8 lanes x 8 B per 64-byte chunk (LDG.64, 4 chunks per warp request)
  // mode 3: 8 lanes x 16 B per 128-byte ALIGNED line (an fp32 head: 4 full lines per warp request)
  constexpr int LPG = MODE == 0 ? 4 : 8;
  const int lane = threadIdx.x % LPG;
  unsigned s = (blockIdx.x * 256 + threadIdx.x / LPG) * 2654435761u + 12345u;
  float acc = 0.f;
  for (int it = 0; it < iters; ++it) {
    uint4 v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int chunk = rng(s) & (window_chunks / 2 - 1);  // power-of-two window: no integer division in the loop
      if (MODE == 2) {
        const uint2 t = SMEM ? reinterpret_cast<const uint2*>(base)[chunk * 8 + lane]
                             : __ldg(reinterpret_cast<const uint2*>(base) + chunk * 8 + lane);
        v[k] = make_uint4(t.x, t.y, t.x, t.y);
      } else if (MODE == 3) {
        v[k] = SMEM ? base[(chunk & ~1) * 4 + lane] : __ldg(base + (chunk & ~1) * 4 + lane);
      } else {
        v[k] = SMEM ? base[chunk * 4 + lane] : __ldg(base + chunk * 4 + lane);
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) acc += __uint_as_float(v[k].x) + __uint_as_float(v[k].w);
  }
  if (acc == 123.456f) *sink = acc;
}

template <int KIND>  // 0: v4.f32, 1: v4.bf16x2
__global__ void __launch_bounds__(256) scatter_kernel(char* dst, size_t n16, int iters, int local) {
  // 4 lanes cover 64 contiguous bytes (bf16) / 8 lanes cover 128 bytes (f32): same shape as the real scatter
  constexpr int LPG = KIND == 0 ? 8 : 4;
  const int lane = threadIdx.x % LPG;
  unsigned s = (blockIdx.x * 256 + threadIdx.x / LPG) * 2654435761u + 777u;
  const size_t groups = n16 / LPG;
  const size_t cta_base = local ? ((size_t)blockIdx.x * 4096) & (groups - 1) : 0;
  for (int it = 0; it < iters; ++it) {
    size_t g = local ? (cta_base + (rng(s) & 2047)) & (groups - 1) : (((size_t)rng(s) << 8) ^ rng(s)) & (groups - 1);
    char* p = dst + (g * LPG + lane) * 16;
    if (KIND == 0)
      asm volatile("red.global.add.v4.f32 [%0], {%1,%1,%1,%1};" ::"l"(p), "f"(1.0f) : "memory");
    else
      asm volatile("red.global.add.noftz.v4.bf16x2 [%0], {%1,%1,%1,%1};" ::"l"(p), "r"(0x3f803f80u) : "memory");
  }
}

// KIND 0: float atomicAdd (CAS loop in SASS), 1: int atomicAdd (native ATOMS.ADD), random addresses
// KIND 2: int atomicAdd, conflict-free (lane i -> bank i)
template <int KIND>
__global__ void __launch_bounds__(256) smem_atomic_kernel(int iters, int words, float* sink) {
  extern __shared__ float smf[];
  int* smi = reinterpret_cast<int*>(smf);
  for (int i = threadIdx.x; i < words; i += blockDim.x) smf[i] = 0.f;
  __syncthreads();
  unsigned s = (blockIdx.x * 256 + threadIdx.x) * 2654435761u + 99u;
  for (int it = 0; it < iters; ++it) {
    const unsigned r = rng(s);
    if (KIND == 0) atomicAdd(&smf[r & (words - 1)], 1.0f);
    else if (KIND == 1) atomicAdd(&smi[r & (words - 1)], 1);
    else atomicAdd(&smi[((r & (words - 1)) & ~31) | (threadIdx.x & 31)], 1);
  }
  __syncthreads();
  if (smf[threadIdx.x] == -1.f) *sink = 1.f;
}

template <typename F>
float time_ms(F f, int reps = 5) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  f(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(a)); f(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  return best;
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  printf("device %s, %d SMs\n", prop.name, sms);
  float* sink; CK(cudaMalloc(&sink, 4));
  // ---- gathers: window of 1024 chunks = 64 KB per CTA, 2 CTAs per SM
  const int window = 1024, iters = 2000, grid = sms * 2;
  uint4* src; CK(cudaMalloc(&src, (size_t)grid * window * 64)); CK(cudaMemset(src, 0, (size_t)grid * window * 64));
  const double bytes = (double)grid * 256 * iters * 4 * 16;
  auto rep = [&](const char* name, float ms) { printf("%-44s %8.3f ms  %8.1f GB/s  (%.1f B/clk/SM @1.965GHz)\n", name, ms, bytes / ms / 1e6, bytes / ms / 1e6 / sms / 1.965); };
  CK(cudaFuncSetAttribute(gather_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, window * 64));
  CK(cudaFuncSetAttribute(gather_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, window * 64));
  rep("LDG.128 64B chunks (L1, 64KB window)", time_ms([&] { gather_kernel<0, false><<<grid, 256>>>(src, window, iters, sink); }));
  rep("LDG.128 128B pairs  (L1, 64KB window)", time_ms([&] { gather_kernel<1, false><<<grid, 256>>>(src, window, iters, sink); }));
  rep("LDS.128 64B chunks (smem 64KB)", time_ms([&] { gather_kernel<0, true><<<grid, 256, window * 64>>>(src, window, iters, sink); }));
  rep("LDS.128 128B pairs  (smem 64KB)", time_ms([&] { gather_kernel<1, true><<<grid, 256, window * 64>>>(src, window, iters, sink); }));
  {
    const double b2 = (double)grid * 256 * iters * 4 * 8;
    float ms = time_ms([&] { gather_kernel<2, false><<<grid, 256>>>(src, window, iters, sink); });
    printf("%-44s %8.3f ms  %8.1f GB/s  (%.1f B/clk/SM @1.965GHz)\n", "LDG.64 64B chunks, 8 lanes (L1, 64KB window)", ms, b2 / ms / 1e6, b2 / ms / 1e6 / sms / 1.965);
  }
  rep("LDG.128 aligned 128B lines (L1, 64KB window)", time_ms([&] { gather_kernel<3, false><<<grid, 256>>>(src, window, iters, sink); }));
  // smaller window: 16 KB
  rep("LDG.128 64B chunks (L1, 16KB window)", time_ms([&] { gather_kernel<0, false><<<grid, 256>>>(src, 256, iters, sink); }));
  // more CTAs per SM (occupancy 8)
  {
    const int g8 = sms * 8;
    uint4* src8; CK(cudaMalloc(&src8, (size_t)g8 * 256 * 64)); CK(cudaMemset(src8, 0, (size_t)g8 * 256 * 64));
    const double b8 = (double)g8 * 256 * iters * 4 * 16;
    float ms = time_ms([&] { gather_kernel<0, false><<<g8, 256>>>(src8, 256, iters, sink); });
    printf("%-44s %8.3f ms  %8.1f GB/s  (%.1f B/clk/SM @1.965GHz)\n", "LDG.128 64B chunks (L1, 16KB win, 8 CTA/SM)", ms, b8 / ms / 1e6, b8 / ms / 1e6 / sms / 1.965);
    ms = time_ms([&] { gather_kernel<1, false><<<g8, 256>>>(src8, 256, iters, sink); });
    printf("%-44s %8.3f ms  %8.1f GB/s  (%.1f B/clk/SM @1.965GHz)\n", "LDG.128 128B pairs  (L1, 16KB win, 8 CTA/SM)", ms, b8 / ms / 1e6, b8 / ms / 1e6 / sms / 1.965);
    CK(cudaFree(src8));
  }
  // ---- L2 gather: window far larger than L1 (whole 88 MB buffer shared by all CTAs)
  {
    const size_t big_chunks = (size_t)88 * 1024 * 1024 / 64;
    uint4* big; CK(cudaMalloc(&big, big_chunks * 64)); CK(cudaMemset(big, 0, big_chunks * 64));
    const int g8 = sms * 8, it2 = 500;
    const double b8 = (double)g8 * 256 * it2 * 4 * 16;
    // every CTA uses base = big (blockIdx * window = 0 by passing a window that wraps): emulate by window=chunks/grid
    float ms = time_ms([&] { gather_kernel<0, false><<<g8, 256>>>(big, (int)(big_chunks / g8), it2, sink); });
    printf("%-44s %8.3f ms  %8.1f GB/s\n", "LDG.128 64B chunks (L2, 74KB window/CTA x8/SM)", ms, b8 / ms / 1e6);
    CK(cudaFree(big));
  }
  // ---- scatters
  for (int big = 0; big < 2; ++big) {
    const size_t nbytes = big ? (size_t)256 * 1024 * 1024 : (size_t)32 * 1024 * 1024;
    char* dst; CK(cudaMalloc(&dst, nbytes)); CK(cudaMemset(dst, 0, nbytes));
    const int g = sms * 8, it = 500;
    const double ops = (double)g * 256 * it;  // lane-level 16-byte reductions
    for (int local = 0; local < 2; ++local) {
      float ms = time_ms([&] { scatter_kernel<0><<<g, 256>>>(dst, nbytes / 16, it, local); });
      printf("red.v4.f32    %3zu MB %-7s %8.3f ms  %7.2f G red16/s  %8.1f GB/s payload\n", nbytes >> 20, local ? "local" : "random", ms, ops / ms / 1e6, ops * 16 / ms / 1e6);
      ms = time_ms([&] { scatter_kernel<1><<<g, 256>>>(dst, nbytes / 16, it, local); });
      printf("red.v4.bf16x2 %3zu MB %-7s %8.3f ms  %7.2f G red16/s  %8.1f GB/s payload\n", nbytes >> 20, local ? "local" : "random", ms, ops / ms / 1e6, ops * 16 / ms / 1e6);
    }
    CK(cudaFree(dst));
  }
  // ---- shared-memory float atomics
  {
    const int g = sms * 4, it = 4000, words = 8192;
    const double ops = (double)g * 256 * it;
    float ms = time_ms([&] { smem_atomic_kernel<0><<<g, 256, words * 4>>>(it, words, sink); });
    printf("smem atomicAdd(float) random 32KB:        %8.3f ms  %7.2f G atom/s  (%.2f lanes/clk/SM @1.965GHz)\n", ms, ops / ms / 1e6, ops / ms / 1e6 / sms / 1.965);
    ms = time_ms([&] { smem_atomic_kernel<1><<<g, 256, words * 4>>>(it, words, sink); });
    printf("smem atomicAdd(int)   random 32KB:        %8.3f ms  %7.2f G atom/s  (%.2f lanes/clk/SM @1.965GHz)\n", ms, ops / ms / 1e6, ops / ms / 1e6 / sms / 1.965);
    ms = time_ms([&] { smem_atomic_kernel<2><<<g, 256, words * 4>>>(it, words, sink); });
    printf("smem atomicAdd(int)   conflict-free:      %8.3f ms  %7.2f G atom/s  (%.2f lanes/clk/SM @1.965GHz)\n", ms, ops / ms / 1e6, ops / ms / 1e6 / sms / 1.965);
  }
  return 0;
}
