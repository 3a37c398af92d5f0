// micro2.cu -- round-2 redo of the load-path microbenchmark (not part of the product).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o micro2 micro2.cu && ./micro2
//
// Round 1's micro.cu was invalid: its 128-bit loads only consumed .x and .w, so the compiler emitted two 32-bit loads
// per "128-bit" access (SASS: LDS R, [R]; LDS R, [R+0xc]) and every conclusion drawn from it ("shared memory is no
// faster than L1", "52 B/clk/SM LSU floor") measured a different access pattern. This version
//   * issues the loads as volatile inline PTX (ld.shared.v4.u32 / ld.global.nc.v4.u32 ...), all four words consumed,
//   * takes its addresses from per-thread offset tables precomputed on the host (no address arithmetic chain),
//   * keeps NOFF = 8 independent loads in flight per thread, at 2 / 4 / 8 blocks of 256 threads per SM,
//   * reports bytes / clk / SM from the SM cycle counter (clock64) as well as from wall time.
// Patterns (W = access width in bytes per lane):
//   chunk64   4 lanes x 16 B (or 8 x 8 B) cover one random 64-byte run; 8 (4) runs per warp request, each in its own line
//   chunk64p  as chunk64 but runs come in pairs sharing one 128-byte line (4 lines per request)
//   chunk64alt  8 runs per request in 8 DIFFERENT lines, lane groups alternating between the lower and the upper half of
//             their line (what msda_fwd_pair_kernel issues: bank-conflict-free, but eight tags per request)
//   chunk64same 8 runs in 8 different lines, all in the lower half (what the one-head forward issues)
//   pair128   8 lanes x 16 B cover 128 contiguous bytes at a random 64-byte-granular offset (two x-adjacent pixels)
//   line128   8 lanes x 16 B cover one random ALIGNED 128-byte line
//   linear    lane i reads base + 16 i (fully coalesced)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <string>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int NT = 256, NOFF = 8;

template <int W, bool SMEM>
__global__ void __launch_bounds__(NT) gather_kernel(const char* __restrict__ src, const int* __restrict__ offs,
                                                    int window_bytes, int iters, unsigned* sink, long long* cycles) {
  extern __shared__ uint4 sm[];
  const char* gbase = src + (size_t)blockIdx.x * window_bytes;
  if (SMEM) {
    for (int i = threadIdx.x; i < window_bytes / 16; i += NT) sm[i] = reinterpret_cast<const uint4*>(gbase)[i];
  }
  int off[NOFF];
#pragma unroll
  for (int k = 0; k < NOFF; ++k) off[k] = offs[k * NT + threadIdx.x];
  const unsigned sbase = (unsigned)__cvta_generic_to_shared(sm);
  unsigned acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    // loop-variant addresses (bits 7..13 are XORed with a hash of the iteration; the bank and line-count shape of every pattern is kept):
    // without this ptxas hoists the loads out of the loop (seen in SASS), volatile asm or not
    const int flip = (int)(((unsigned)it * 0x9E3779B1u) >> 25) << 7;  // bits 7..13: same banks, same lines-per-request
#pragma unroll
    for (int k = 0; k < NOFF; ++k) {
      unsigned x = 0, y = 0, z = 0, w = 0;
      if (SMEM) {
        const unsigned a = sbase + (off[k] ^ flip);
        if (W == 16) asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "r"(a));
        if (W == 8) asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(x), "=r"(y) : "r"(a));
        if (W == 4) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(x) : "r"(a));
      } else {
        const char* a = gbase + (off[k] ^ flip);
        if (W == 16) asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "l"(a));
        if (W == 8) asm volatile("ld.global.nc.v2.u32 {%0,%1}, [%2];" : "=r"(x), "=r"(y) : "l"(a));
        if (W == 4) asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(x) : "l"(a));
      }
      acc ^= x ^ y;
      acc ^= z ^ w;
    }
  }
  const long long t1 = clock64();
  if (acc == 0x12345678u) *sink = acc;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// shared-memory atomics and shuffles, for the backward design
template <int KIND>  // 0: ATOMS.ADD int random, 1: ATOMS.ADD with return value (used by a counting-sort fill), 2: SHFL.BFLY
__global__ void __launch_bounds__(NT) misc_kernel(int iters, int words, unsigned* sink, long long* cycles) {
  extern __shared__ int smi[];
  for (int i = threadIdx.x; i < words; i += NT) smi[i] = 0;
  __syncthreads();
  unsigned s = (blockIdx.x * NT + threadIdx.x) * 2654435761u + 99u;
  unsigned acc = 0;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    s = s * 1664525u + 1013904223u;
    const unsigned r = (s >> 8) & (words - 1);
    if (KIND == 0) atomicAdd(&smi[r], (int)(s >> 28) + 1);
    if (KIND == 1) acc += atomicAdd(&smi[r], 1);
    if (KIND == 2) acc += __shfl_xor_sync(0xffffffffu, s, 4);
  }
  const long long t1 = clock64();
  __syncthreads();
  if (acc == 0x12345678u || smi[threadIdx.x] == -1) *sink = acc;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

static unsigned lcg(unsigned& s) { s = s * 1664525u + 1013904223u; return s >> 8; }

enum Pattern { CHUNK64, CHUNK64P, PAIR128, LINE128, LINEAR, CHUNK64ALT, CHUNK64SAME };

static std::vector<int> make_offsets(Pattern pat, int W, int window) {
  std::vector<int> o(NOFF * NT);
  unsigned s = 12345u;
  const int lanes64 = 64 / W, lanes128 = 128 / W;
  for (int k = 0; k < NOFF; ++k) {
    std::vector<int> grp(NT);
    for (int g = 0; g < NT; ++g) grp[g] = (int)lcg(s);
    for (int t = 0; t < NT; ++t) {
      int v = 0;
      switch (pat) {
        case CHUNK64: v = (grp[t / lanes64] % (window / 64)) * 64 + (t % lanes64) * W; break;
        case CHUNK64P: v = (grp[t / lanes128] % (window / 128)) * 128 + (t % lanes128) * W; break;  // == LINE128 addresses
        case PAIR128: v = (grp[t / lanes128] % (window / 64 - 1)) * 64 + (t % lanes128) * W; break;
        case LINE128: v = (grp[t / lanes128] % (window / 128)) * 128 + (t % lanes128) * W; break;
        case LINEAR: v = ((k * NT + t) * W) % window; break;
        case CHUNK64ALT:
        case CHUNK64SAME: {
          const int gi = t / lanes64, per_req = 32 / lanes64, j = gi % per_req, lines = window / 128;
          const int line = (grp[gi / per_req] + j * (lines / per_req)) % lines;  // per_req distinct lines per request
          v = line * 128 + (pat == CHUNK64ALT ? (j & 1) * 64 : 0) + (t % lanes64) * W;
          break;
        }
      }
      o[k * NT + t] = v;
    }
  }
  return o;
}

int main(int argc, char** argv) {
  const bool quick = argc > 1;  // `micro2 ncu`: one launch per pattern at occ 4, for an ncu capture
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  int khz = 0; CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
  printf("device %s, %d SMs, max clock %d MHz\n", prop.name, sms, khz / 1000);
  unsigned* sink; CK(cudaMalloc(&sink, 4));
  const int window = 16 * 1024, iters = quick ? 500 : 4000;
  const int max_grid = sms * 8;
  char* src; CK(cudaMalloc(&src, (size_t)max_grid * window)); CK(cudaMemset(src, 1, (size_t)max_grid * window));
  int* d_off; CK(cudaMalloc(&d_off, sizeof(int) * NOFF * NT));
  long long* d_cyc; CK(cudaMalloc(&d_cyc, sizeof(long long) * max_grid));
  std::vector<long long> h_cyc(max_grid);
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));

  auto run = [&](const char* name, auto kern, bool smem, int W, Pattern pat, int occ) {
    auto o = make_offsets(pat, W, window);
    CK(cudaMemcpy(d_off, o.data(), sizeof(int) * o.size(), cudaMemcpyHostToDevice));
    const int grid = sms * occ;
    // occupancy is pinned by the dynamic shared memory request: 227 KB / occ (L1 variants still get >= 64 KB of L1 at occ <= 8... the
    // request only limits residency, the kernel uses `window` bytes of it when SMEM)
    const int smem_bytes = smem ? ((200 * 1024 / occ) / 1024) * 1024 : 0;
    if (smem) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    float best = 1e30f;
    for (int r = 0; r < (quick ? 1 : 4); ++r) {
      CK(cudaEventRecord(e0));
      kern<<<grid, NT, smem_bytes>>>(src, d_off, window, iters, sink, d_cyc);
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      if ((quick || r > 0) && ms < best) best = ms;
    }
    CK(cudaGetLastError());
    CK(cudaMemcpy(h_cyc.data(), d_cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
    double cyc = 0; for (int i = 0; i < grid; ++i) cyc += (double)h_cyc[i];
    cyc /= grid;  // mean cycles a block was resident; occ blocks share the SM for that long
    const double bytes_per_block = (double)NT * iters * NOFF * W;
    printf("%-34s W=%2d occ=%d  %8.3f ms  %7.1f GB/s  %6.1f B/clk/SM (clock64)  %6.1f (wall @%d MHz)\n", name, W, occ, best,
           bytes_per_block * grid / best / 1e6, bytes_per_block * occ / cyc,
           bytes_per_block * grid / best / 1e6 / sms / (khz / 1e6), khz / 1000);
  };

  for (int occ : {2, 4, 8}) {
    if (quick && occ != 4) continue;
    run("LDS chunk64", gather_kernel<16, true>, true, 16, CHUNK64, occ);
    run("LDS pair128", gather_kernel<16, true>, true, 16, PAIR128, occ);
    run("LDS line128", gather_kernel<16, true>, true, 16, LINE128, occ);
    run("LDS linear", gather_kernel<16, true>, true, 16, LINEAR, occ);
    run("LDS chunk64 (8 lanes x 8 B)", gather_kernel<8, true>, true, 8, CHUNK64, occ);
    run("LDS pair128 (16 lanes x 8 B)", gather_kernel<8, true>, true, 8, PAIR128, occ);
    run("LDS linear", gather_kernel<8, true>, true, 8, LINEAR, occ);
    run("LDS linear", gather_kernel<4, true>, true, 4, LINEAR, occ);
    run("LDG chunk64 (8 lines/request)", gather_kernel<16, false>, false, 16, CHUNK64, occ);
    run("LDG chunk64p (4 lines/request)", gather_kernel<16, false>, false, 16, CHUNK64P, occ);
    run("LDG chunk64alt (8 lines, alt. halves)", gather_kernel<16, false>, false, 16, CHUNK64ALT, occ);
    run("LDG chunk64same (8 lines, one half)", gather_kernel<16, false>, false, 16, CHUNK64SAME, occ);
    run("LDS chunk64alt (8 lines, alt. halves)", gather_kernel<16, true>, true, 16, CHUNK64ALT, occ);
    run("LDG pair128 (4-8 lines/request)", gather_kernel<16, false>, false, 16, PAIR128, occ);
    run("LDG line128", gather_kernel<16, false>, false, 16, LINE128, occ);
    run("LDG linear", gather_kernel<16, false>, false, 16, LINEAR, occ);
    run("LDG chunk64 (8 lanes x 8 B)", gather_kernel<8, false>, false, 8, CHUNK64, occ);
    run("LDG pair128 (16 lanes x 8 B)", gather_kernel<8, false>, false, 8, PAIR128, occ);
    run("LDG linear", gather_kernel<8, false>, false, 8, LINEAR, occ);
    run("LDG linear", gather_kernel<4, false>, false, 4, LINEAR, occ);
  }

  auto run_misc = [&](const char* name, auto kern, int occ, int words) {
    const int grid = sms * occ;
    const int it2 = 20000;
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
      CK(cudaEventRecord(e0));
      kern<<<grid, NT, words * 4>>>(it2, words, sink, d_cyc);
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      if (r > 0 && ms < best) best = ms;
    }
    CK(cudaGetLastError());
    CK(cudaMemcpy(h_cyc.data(), d_cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
    double cyc = 0; for (int i = 0; i < grid; ++i) cyc += (double)h_cyc[i];
    cyc /= grid;
    printf("%-34s occ=%d words=%d  %8.3f ms  %6.2f lanes/clk/SM (clock64)\n", name, occ, words, best, (double)NT * it2 * occ / cyc);
  };
  for (int occ : {4, 8}) {
    if (quick) break;
    run_misc("ATOMS.ADD (no return) random", misc_kernel<0>, occ, 2048);
    run_misc("ATOMS.ADD (returning) random", misc_kernel<1>, occ, 2048);
    run_misc("SHFL.BFLY", misc_kernel<2>, occ, 2048);
  }
  printf("done\n");
  return 0;
}
