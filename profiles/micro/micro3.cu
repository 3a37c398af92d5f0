// micro3.cu -- ldmatrix / mma.sync throughput on B200 (round 2, sizing the tensor-core gather of csrc/msda_win.cu).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o micro3 micro3.cu && ./micro3
// Each kernel runs `iters` iterations of 8 independent operations per warp, 8 warps per block, occ blocks per SM,
// and reports operations per clock per SM from clock64().
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
constexpr int NT = 256;

// MODE 0: ldmatrix.x4 (rows = 16-byte chunks at conflict-free addresses), 1: ldmatrix.x4.trans, 2: x4 with all 8 rows of a
// matrix in the same bank group (the 4-way conflict the rotation trick removes), 3: mma.sync m16n8k16 bf16 (8 independent
// accumulators), 4: mma.sync in one dependent chain (latency)
template <int MODE>
__global__ void __launch_bounds__(NT) k(int iters, unsigned* sink, long long* cycles) {
  extern __shared__ __align__(128) unsigned char sm[];
  for (int i = threadIdx.x; i < 16384 / 4; i += NT) reinterpret_cast<unsigned*>(sm)[i] = i * 2654435761u;
  __syncthreads();
  const unsigned base = (unsigned)__cvta_generic_to_shared(sm);
  const int lane = threadIdx.x & 31;
  // row address of this lane: matrix (lane>>3), row (lane&7)
  unsigned addr[8];
  for (int j = 0; j < 8; ++j) {
    const int m = lane >> 3, r = lane & 7;
    if (MODE == 2) addr[j] = base + ((j * 4 + m) * 8 + r) * 128 % 16384;          // same 16-byte slot in every row
    else addr[j] = base + (((j * 4 + m) * 8 + r) * 144) % 16256 / 16 * 16;         // stride 144 B: slots rotate
  }
  unsigned acc = 0;
  float d[8][4];
  for (int j = 0; j < 8; ++j) for (int q = 0; q < 4; ++q) d[j][q] = 0.f;
  unsigned a0 = lane * 7u + 1, a1 = lane * 5u + 3, a2 = lane * 3u + 7, a3 = lane + 11u, b0 = lane * 13u, b1 = lane * 17u;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const unsigned flip = ((unsigned)it * 0x9E3779B1u >> 25) << 7;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (MODE <= 2) {
        unsigned r0, r1, r2, r3;
        const unsigned a = base + ((addr[j] - base) ^ flip);
        if (MODE == 1) asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(a));
        else asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(a));
        acc ^= r0 ^ r1;
        acc ^= r2 ^ r3;
      } else {
        const int jj = MODE == 3 ? j : 0;
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(d[jj][0]), "+f"(d[jj][1]), "+f"(d[jj][2]), "+f"(d[jj][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
      }
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
  for (int j = 0; j < 8; ++j) for (int q = 0; q < 4; ++q) s += d[j][q];
  if (acc == 0x12345678u || s == 1.2345f) *sink = acc;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  unsigned* sink; CK(cudaMalloc(&sink, 4));
  long long* d_cyc; CK(cudaMalloc(&d_cyc, sizeof(long long) * sms * 8));
  std::vector<long long> h(sms * 8);
  const int iters = 2000;
  auto run = [&](const char* name, auto kern, int occ) {
    const int grid = sms * occ;
    const int smem = ((200 * 1024 / occ) / 1024) * 1024;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
      CK(cudaEventRecord(e0)); kern<<<grid, NT, smem>>>(iters, sink, d_cyc); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    const double ops = (double)grid * (NT / 32) * iters * 8;  // warp-level operations
    printf("%-44s occ=%d %8.3f ms  %6.3f warp-ops/clk/SM (wall @1965 MHz)  = %5.2f clk per op per SM\n", name, occ, best,
           ops / (best * 1e-3) / sms / 1.965e9, (best * 1e-3) * sms * 1.965e9 / ops);
  };
  for (int occ : {2, 4}) {
    run("ldmatrix.x4 conflict-free", k<0>, occ);
    run("ldmatrix.x4.trans conflict-free", k<1>, occ);
    run("ldmatrix.x4 same slot in all rows (8-way)", k<2>, occ);
    run("mma.sync m16n8k16 bf16, independent", k<3>, occ);
    run("mma.sync m16n8k16 bf16, dependent chain", k<4>, occ);
  }
  return 0;
}
