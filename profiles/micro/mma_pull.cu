// mma_pull.cu -- PROTOTYPE (not part of the product, NOT yet run on hardware: the GPU budget of round 1 was spent
// when it was written; it compiles for sm_100a and self-checks when run).  Static result so far (cuobjdump -sass):
// the MMA loop is ~200 instructions per 16 entries (~170 when the chunk has <= 8 distinct pixels) against ~150 for the
// CUDA-core loop on the same 16 entries, so it is not expected to win as written -- see DESIGN.md section 6.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o mma_pull mma_pull.cu && ./mma_pull
//
// Question it answers: is the "pull" of the pixel-sorted backward (csrc/msda_bwd_sorted.cuh, step f) cheaper on
// warp-level tensor-core MMAs than on CUDA cores?  Both kernels below consume the same per-block inputs the real
// kernel has in shared memory after its counting sort:
//   ent[E]      sorted entries {pixel << 16 | id, weight}; id >> 4 = row of the staged grad_out tile (query)
//   go[128][32] the tile's grad_out rows for one head, bf16
//   V[pix][32]  value rows, bf16, global memory
// and produce   dots[id] = <go[row(id)], V[pixel]>            (feeds grad_attn / grad_loc)
//               acc[pix][32] += weight * go[row(id)]           (grad_value, fp32 global reductions)
//
// pull_scalar: the product's loop (4 lanes per entry, equal shares, register accumulator per run, FFMA).
// pull_mma   : a warp takes chunks of 16 consecutive entries.  Per chunk:
//   * slots = distinct pixels of the chunk (ballot of "new pixel" flags, popc ranks);
//   * acc:  C[slot, ch] = A[slot, entry] * GF[entry, ch]   m16n8k16, A one-hot-per-entry weights split into bf16
//           hi + lo (two MMAs: products exact to 2^-17), GF fragments by ldmatrix.trans with per-entry row addresses
//           (no bf16 -> fp32 unpack);
//   * dots: D[entry, slot] = GF[entry, ch] * V[slot, ch]^T  (ldmatrix, V fragments by 4-byte global loads);
//           lane picks D[entry, slot(entry)];
//   * every chunk flushes its slots with red.global.add.v2.f32 (runs crossing a chunk boundary cost one extra
//     reduction; reductions are not the bottleneck).
// The weight split is done when the entry is written (here: on the host), so the pull only selects.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

constexpr int NT = 256;        // threads per block
constexpr int E = 2048;        // entries per block (one level of one 128-query tile)
constexpr int TQ = 128;        // grad_out rows in the tile
constexpr int NPIX = 4096;     // value rows
constexpr int D = 32;          // channels per head

struct Smem {
  int2 ent[E];
  uint4 go[TQ * 4];            // 64 bytes per row
  int slotpix[NT / 32][16];
};

__device__ __forceinline__ void red_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red_v2(float* p, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void unpack8(uint4 v, float* f) {
  f[0] = __uint_as_float(v.x << 16); f[1] = __uint_as_float(v.x & 0xffff0000u);
  f[2] = __uint_as_float(v.y << 16); f[3] = __uint_as_float(v.y & 0xffff0000u);
  f[4] = __uint_as_float(v.z << 16); f[5] = __uint_as_float(v.z & 0xffff0000u);
  f[6] = __uint_as_float(v.w << 16); f[7] = __uint_as_float(v.w & 0xffff0000u);
}

__device__ __forceinline__ void stage(Smem& s, const int2* ent, const uint4* go) {
  for (int i = threadIdx.x; i < E; i += NT) s.ent[i] = ent[(size_t)blockIdx.x * E + i];
  for (int i = threadIdx.x; i < TQ * 4; i += NT) s.go[i] = go[(size_t)blockIdx.x * TQ * 4 + i];
  __syncthreads();
}

// ------------------------------------------------------------------------------------------- scalar (product loop)
__global__ void __launch_bounds__(NT, 4) pull_scalar(const int2* __restrict__ ent, const uint4* __restrict__ go,
                                                     const uint4* __restrict__ V, float* __restrict__ acc_out,
                                                     float* __restrict__ dots) {
  extern __shared__ __align__(16) unsigned char raw[];
  Smem& s = *reinterpret_cast<Smem*>(raw);
  stage(s, ent, go);
  const int tid = threadIdx.x, pg = tid >> 2, pc = tid & 3;
  constexpr int GP = NT / 4;
  const int per = ((E + GP - 1) / GP) | 1;
  const int e0 = pg * per, e_end = min(e0 + per, E);
  int cur = -1;
  float acc[8], vf[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { acc[j] = 0.f; vf[j] = 0.f; }
  float* dot_base = dots + (size_t)blockIdx.x * E;
  for (int k = 0; k < per; ++k) {
    const int e = e0 + k;
    const bool valid = e < e_end;
    int2 en = make_int2(cur << 16, 0);
    if (valid) en = s.ent[e];
    const int pix = (int)((unsigned)en.x >> 16), id = en.x & 0xffff;
    const float w = __int_as_float(en.y);
    if (valid && pix != cur) {
      if (cur >= 0) {
        float* dst = acc_out + (size_t)cur * D + pc * 8;
        red_v4(dst, acc[0], acc[1], acc[2], acc[3]);
        red_v4(dst + 4, acc[4], acc[5], acc[6], acc[7]);
      }
      cur = pix;
      unpack8(__ldg(V + (size_t)pix * 4 + pc), vf);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    }
    float gf[8];
    unpack8(s.go[(id >> 4) * 4 + pc], gf);
    float d = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) { acc[j] = fmaf(w, gf[j], acc[j]); d = fmaf(gf[j], vf[j], d); }
    d += __shfl_xor_sync(0xffffffffu, d, 1);
    d += __shfl_xor_sync(0xffffffffu, d, 2);
    if (valid && pc == 0) dot_base[id] = d;
  }
  if (cur >= 0) {
    float* dst = acc_out + (size_t)cur * D + pc * 8;
    red_v4(dst, acc[0], acc[1], acc[2], acc[3]);
    red_v4(dst + 4, acc[4], acc[5], acc[6], acc[7]);
  }
}

// ------------------------------------------------------------------------------------------------- warp MMA variant
__device__ __forceinline__ void ldsm_x4(unsigned addr, unsigned& r0, unsigned& r1, unsigned& r2, unsigned& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(unsigned addr, unsigned& r0, unsigned& r1, unsigned& r2, unsigned& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const unsigned (&a)[4], unsigned b0, unsigned b1,
                                         const float (&c)[4]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
               "{%10, %11, %12, %13};"
               : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(c[0]), "f"(c[1]), "f"(c[2]), "f"(c[3]));
}

// ent.y for this kernel: bf16 hi in the low half, bf16 lo in the high half (weight = hi + lo up to 2^-17 relative)
__global__ void __launch_bounds__(NT, 4) pull_mma(const int2* __restrict__ ent, const uint4* __restrict__ go,
                                                  const unsigned* __restrict__ V32, float* __restrict__ acc_out,
                                                  float* __restrict__ dots) {
  extern __shared__ __align__(16) unsigned char raw[];
  Smem& s = *reinterpret_cast<Smem*>(raw);
  stage(s, ent, go);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const unsigned go_base = (unsigned)__cvta_generic_to_shared(s.go);
  float* dot_base = dots + (size_t)blockIdx.x * E;
  int* slotpix = s.slotpix[warp];
  const float zero4[4] = {0.f, 0.f, 0.f, 0.f};
  constexpr int CHUNKS = E / 16, WARPS = NT / 32;
  for (int ch = warp; ch < CHUNKS; ch += WARPS) {   // interleaved: balances the warps; any split works
    const int cb = ch * 16;
    // --- own entry (lane & 15): pixel, id, "new pixel" flag, slot ranks
    const int2 own = s.ent[cb + (lane & 15)];
    const int pix = (int)((unsigned)own.x >> 16), id = own.x & 0xffff;
    const int prev = __shfl_up_sync(0xffffffffu, pix, 1);
    const bool flag = ((lane & 15) == 0) || (pix != prev);
    const unsigned mask = __ballot_sync(0xffffffffu, flag) & 0xffffu;
    const int ns = __popc(mask);
    auto slot_of = [&](int e) { return __popc(mask & ((2u << e) - 1u)) - 1; };
    if (lane < 16 && flag) slotpix[slot_of(lane)] = pix;
    __syncwarp();
    // --- A fragments of the weight matrix [slot][entry]: entries 2t, 2t+1, 2t+8, 2t+9; rows g and g+8
    const int4 e01 = *reinterpret_cast<const int4*>(&s.ent[cb + 2 * t]);      // {key, w} of 2t and 2t+1
    const int4 e89 = *reinterpret_cast<const int4*>(&s.ent[cb + 2 * t + 8]);
    const int s0 = slot_of(2 * t), s1 = slot_of(2 * t + 1), s8 = slot_of(2 * t + 8), s9 = slot_of(2 * t + 9);
    const unsigned w0 = (unsigned)e01.y, w1 = (unsigned)e01.w, w8 = (unsigned)e89.y, w9 = (unsigned)e89.w;
    auto pick = [](bool lo_sel, unsigned wlo_idx, bool hi_sel, unsigned whi_idx, int shift) -> unsigned {
      const unsigned a = lo_sel ? ((wlo_idx >> shift) & 0xffffu) : 0u;
      const unsigned b = hi_sel ? ((whi_idx >> shift) & 0xffffu) : 0u;
      return a | (b << 16);
    };
    unsigned a_hi[4], a_lo[4];
    a_hi[0] = pick(s0 == g, w0, s1 == g, w1, 0);          a_lo[0] = pick(s0 == g, w0, s1 == g, w1, 16);
    a_hi[1] = pick(s0 == g + 8, w0, s1 == g + 8, w1, 0);  a_lo[1] = pick(s0 == g + 8, w0, s1 == g + 8, w1, 16);
    a_hi[2] = pick(s8 == g, w8, s9 == g, w9, 0);          a_lo[2] = pick(s8 == g, w8, s9 == g, w9, 16);
    a_hi[3] = pick(s8 == g + 8, w8, s9 == g + 8, w9, 0);  a_lo[3] = pick(s8 == g + 8, w8, s9 == g + 8, w9, 16);
    // --- grad_out fragments: lane's own entry gives the row; matrices (lane>>3): entries 0-7 / 8-15 x channel tiles
    const unsigned row_addr = go_base + (unsigned)(id >> 4) * 64u + (unsigned)(lane >> 4) * 16u;
    unsigned bt[2][4], an[2][4];
    ldsm_x4_t(row_addr, bt[0][0], bt[0][1], bt[0][2], bt[0][3]);        // channel tiles 0, 1 (B operand, k = entry)
    ldsm_x4_t(row_addr + 32u, bt[1][0], bt[1][1], bt[1][2], bt[1][3]);  // channel tiles 2, 3
    ldsm_x4(row_addr, an[0][0], an[0][1], an[0][2], an[0][3]);          // A operand of the dots, k = channels 0-15
    ldsm_x4(row_addr + 32u, an[1][0], an[1][1], an[1][2], an[1][3]);    // channels 16-31
    // --- grad_value: C[slot][ch] over four channel tiles
    float c[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      mma_bf16(c[j], a_hi, bt[j >> 1][(j & 1) * 2], bt[j >> 1][(j & 1) * 2 + 1], zero4);
      mma_bf16(c[j], a_lo, bt[j >> 1][(j & 1) * 2], bt[j >> 1][(j & 1) * 2 + 1], c[j]);
    }
    // --- dots: D[entry][slot] over up to two slot tiles; B operand = V[slot g (+8)][channels], 4-byte loads
    const int sg = slot_of(g), sg8 = slot_of(g + 8);                 // slots of the entries in rows g, g+8
    const int id_g = __shfl_sync(0xffffffffu, id, g), id_g8 = __shfl_sync(0xffffffffu, id, g + 8);
    float dot_g = 0.f, dot_g8 = 0.f;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (u * 8 < ns) {  // warp-uniform
        const int slot = u * 8 + g;
        const unsigned* vrow = V32 + (size_t)slotpix[min(slot, ns - 1)] * 16;  // clamp: unused columns read a valid row
        const unsigned b0 = __ldg(vrow + t), b1 = __ldg(vrow + 4 + t), b2 = __ldg(vrow + 8 + t), b3 = __ldg(vrow + 12 + t);
        float dd[4];
        mma_bf16(dd, an[0], b0, b1, zero4);
        mma_bf16(dd, an[1], b2, b3, dd);
        // dd[0..1] = D[g][slots u*8 + 2t, +1], dd[2..3] = D[g+8][same]
        if ((sg >> 3) == u && ((sg & 7) >> 1) == t) dot_g = (sg & 1) ? dd[1] : dd[0];
        if ((sg8 >> 3) == u && ((sg8 & 7) >> 1) == t) dot_g8 = (sg8 & 1) ? dd[3] : dd[2];
      }
    }
    if (((sg & 7) >> 1) == t) dot_base[id_g] = dot_g;
    if (((sg8 & 7) >> 1) == t) dot_base[id_g8] = dot_g8;
    // --- flush the chunk's slots: rows g and g+8 of the four channel tiles, two channels per lane
    if (g < ns) {
      float* dst = acc_out + (size_t)slotpix[g] * D + 2 * t;
#pragma unroll
      for (int j = 0; j < 4; ++j) red_v2(dst + 8 * j, c[j][0], c[j][1]);
    }
    if (g + 8 < ns) {
      float* dst = acc_out + (size_t)slotpix[g + 8] * D + 2 * t;
#pragma unroll
      for (int j = 0; j < 4; ++j) red_v2(dst + 8 * j, c[j][2], c[j][3]);
    }
    __syncwarp();  // slotpix is rewritten by the next chunk
  }
}

// ------------------------------------------------------------------------------------------------------------ host
static unsigned short f2bf(float f) {
  unsigned u;
  memcpy(&u, &f, 4);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (unsigned short)(u >> 16);
}
static float bf2f(unsigned short h) {
  unsigned u = (unsigned)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

int main() {
  int dev = 0;
  CK(cudaSetDevice(dev));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, dev));
  const int NB = prop.multiProcessorCount * 4 * 8;
  printf("device %s, %d SMs, %d blocks of %d entries\n", prop.name, prop.multiProcessorCount, NB, E);
  std::vector<int2> ent_f((size_t)NB * E), ent_s((size_t)NB * E);
  std::vector<unsigned short> go((size_t)NB * TQ * D), V((size_t)NPIX * D);
  unsigned rs = 12345u;
  auto rnd = [&]() { rs = rs * 1664525u + 1013904223u; return rs >> 8; };
  auto rndf = [&]() { return (float)(rnd() & 0xffff) / 32768.f - 1.f; };
  for (auto& v : go) v = f2bf(rndf());
  for (auto& v : V) v = f2bf(rndf());
  double mean_run = 0;
  size_t runs = 0;
  for (int b = 0; b < NB; ++b) {
    int e = 0, pix = rnd() % (NPIX / 2);
    while (e < E) {
      int len = 1 + (int)(rnd() % 13);  // 1..13, mean 7 (the product's measured mean run is 6.9)
      if (e + len > E) len = E - e;
      for (int k = 0; k < len; ++k, ++e) {
        const int id = e;  // ids are a permutation in the product; identity keeps the check simple
        const float w = 0.05f + 0.9f * (float)(rnd() & 0xffff) / 65536.f;
        const unsigned short hi = f2bf(w), lo = f2bf(w - bf2f(hi));
        int wi;
        memcpy(&wi, &w, 4);
        ent_f[(size_t)b * E + e] = make_int2((pix << 16) | id, wi);
        ent_s[(size_t)b * E + e] = make_int2((pix << 16) | id, (int)((unsigned)hi | ((unsigned)lo << 16)));
      }
      ++runs;
      mean_run += len;
      pix = (pix + 1 + (int)(rnd() % 3)) % NPIX;
    }
  }
  printf("mean run length %.2f\n", mean_run / runs);
  int2 *d_ef, *d_es;
  uint4* d_go;
  uint4* d_V;
  float *d_accA, *d_accB, *d_dotA, *d_dotB;
  CK(cudaMalloc(&d_ef, ent_f.size() * sizeof(int2)));
  CK(cudaMalloc(&d_es, ent_s.size() * sizeof(int2)));
  CK(cudaMalloc(&d_go, go.size() * 2));
  CK(cudaMalloc(&d_V, V.size() * 2));
  CK(cudaMalloc(&d_accA, (size_t)NPIX * D * 4));
  CK(cudaMalloc(&d_accB, (size_t)NPIX * D * 4));
  CK(cudaMalloc(&d_dotA, (size_t)NB * E * 4));
  CK(cudaMalloc(&d_dotB, (size_t)NB * E * 4));
  CK(cudaMemcpy(d_ef, ent_f.data(), ent_f.size() * sizeof(int2), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_es, ent_s.data(), ent_s.size() * sizeof(int2), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_go, go.data(), go.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_V, V.data(), V.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaFuncSetAttribute(pull_scalar, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
  CK(cudaFuncSetAttribute(pull_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float msA = 0, msB = 0;
  for (int rep = 0; rep < 6; ++rep) {
    CK(cudaMemset(d_accA, 0, (size_t)NPIX * D * 4));
    CK(cudaMemset(d_accB, 0, (size_t)NPIX * D * 4));
    CK(cudaEventRecord(e0));
    pull_scalar<<<NB, NT, sizeof(Smem)>>>(d_ef, d_go, d_V, d_accA, d_dotA);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaGetLastError());
    float a;
    CK(cudaEventElapsedTime(&a, e0, e1));
    CK(cudaEventRecord(e0));
    pull_mma<<<NB, NT, sizeof(Smem)>>>(d_es, d_go, reinterpret_cast<const unsigned*>(d_V), d_accB, d_dotB);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaGetLastError());
    float bms;
    CK(cudaEventElapsedTime(&bms, e0, e1));
    if (rep >= 2) { msA += a / 4; msB += bms / 4; }
  }
  std::vector<float> accA((size_t)NPIX * D), accB((size_t)NPIX * D), dotA((size_t)NB * E), dotB((size_t)NB * E);
  CK(cudaMemcpy(accA.data(), d_accA, accA.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(accB.data(), d_accB, accB.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(dotA.data(), d_dotA, dotA.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(dotB.data(), d_dotB, dotB.size() * 4, cudaMemcpyDeviceToHost));
  double max_acc = 0, err_acc = 0, max_dot = 0, err_dot = 0;
  for (size_t i = 0; i < accA.size(); ++i) { max_acc = fmax(max_acc, fabs(accA[i])); err_acc = fmax(err_acc, fabs(accA[i] - accB[i])); }
  for (size_t i = 0; i < dotA.size(); ++i) { max_dot = fmax(max_dot, fabs(dotA[i])); err_dot = fmax(err_dot, fabs(dotA[i] - dotB[i])); }
  // reference for block 0 on the host (guards against both kernels being wrong the same way)
  double err_ref = 0;
  for (int e = 0; e < E; ++e) {
    const int2 en = ent_f[e];
    const int pix = (unsigned)en.x >> 16, id = en.x & 0xffff;
    double d = 0;
    for (int j = 0; j < D; ++j) d += (double)bf2f(go[(size_t)(id >> 4) * D + j]) * bf2f(V[(size_t)pix * D + j]);
    err_ref = fmax(err_ref, fabs(d - dotA[e]));
  }
  printf("scalar pull %.3f ms, mma pull %.3f ms (%d blocks; the product's backward runs 32 256 block-levels)\n", msA, msB, NB);
  printf("acc: max %.3f, |scalar - mma| %.3e (rel %.2e)\n", max_acc, err_acc, err_acc / max_acc);
  printf("dot: max %.3f, |scalar - mma| %.3e (rel %.2e); scalar vs host reference (block 0) %.3e\n", max_dot, err_dot,
         err_dot / max_dot, err_ref);
  const bool ok = err_acc / max_acc < 1e-4 && err_dot / max_dot < 1e-5 && err_ref < 1e-3;
  printf(ok ? "CHECK OK\n" : "CHECK FAILED\n");
  return ok ? 0 : 1;
}
