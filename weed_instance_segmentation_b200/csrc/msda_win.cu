// msda_win.cu -- window-staged MSDeformAttn kernels for sm_100a (round 2).
//
// Why: the value gather moves 48 corners x 64 B per (query, head) through the load/store unit. Measured on B200
// (profiles/r02_micro2.log, profiles/micro/micro2.cu): LDS.128 and L1-hit LDG.128 both deliver 127 B/clk/SM when every
// quarter-warp reads 128 CONTIGUOUS bytes and only 82 B/clk/SM on independent 64-byte runs (data-bank conflicts, in
// the L1 data array exactly as in shared memory). In the reference layout (B,S,H,D) two x-adjacent pixels of one head
// are 512 B apart, so the per-corner gather of msda_b200.cu can never see the fast shape. Here a thread block stages
// the bilinear footprint of its query tile -- per level a (bh x bw)-pixel box of ONE head -- in shared memory with TMA
// (cp.async.bulk.tensor, 5-D tensor map over (D, H, W_l, H_l, B), box (32, 1, bw, bh, 1)): inside the box the two
// x-adjacent pixels of a head ARE contiguous (64 B each), so eight lanes read the top (then the bottom) pixel pair of
// a sample as one conflict-free 128-byte run. TMA zero-fills the part of a box that lies outside the level, which is
// exactly grid_sample's padding_mode="zeros" (M2F:823): no clamping, no weight re-slotting, and a non-finite value in
// a pixel the reference never reads cannot leak into the result.
//
// Tiles: the host hands queries out in "pyramid" tiles (functional.pyramid_schedule): an 8x16 patch of the finest level
// plus the coarser-level queries whose reference point falls into that patch, so all queries of a tile look at the
// same small windows. The window ORIGIN is chosen per block from the bounding box of the tile's actual samples; a
// sample whose footprint is not inside the box takes a per-corner global-memory path inside the same loop (any input
// is handled, only slower).
//
// Contract covered: bf16 values, D = 32 (64-byte rows), L <= 4, P = 4; everything else stays on msda_b200.cu.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "msda_b200.h"

extern "C" int msda_b200_internal_fail(int code, const char* msg);

namespace {

#include "msda_common.cuh"

constexpr int kWinL = 4;           // levels a window kernel handles
constexpr int kWinNT = 256;        // threads per block
constexpr int kWinTQ = 192;        // queries per tile (capacity); 8x16 + 4x8 + 2x4 = 168 for a 2x pyramid
constexpr int kWinP = 4;
constexpr int kRowB = 64;          // bytes of one head of one pixel (D = 32, bf16)
constexpr unsigned kFlagSlow = 0x80000000u, kFlagDead = 0x40000000u;

struct alignas(64) WinMaps {
  CUtensorMap m[kWinL];
};

struct WParams {
  KParams k;
  const int* tile_start;  // [num_tiles + 1] offsets into k.q_order
  int bw[kWinL], bh[kWinL];
  int woff[kWinL];        // byte offset of each level's window inside the window area (multiple of 128)
  int win_bytes;
  int guard_bytes;        // zero rows for samples without a window: max pitch + 128
  int tqs;                // descriptor row stride (MMA kernel): max_tile rounded up to 8, plus 1
  int hpb;                // heads per block (MMA kernel): divides H
  long long* dbg;         // optional phase timestamps (dev tool), NULL in production
};

// ---------------------------------------------------------------------------------------------------------------
// mbarrier / TMA / cp.async primitives (PTX; SASS: SYNCS.*, UTMALDG, LDGSTS)
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a tensor map or coordinate bug must end in a launch failure, never in a hung GPU.
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  for (unsigned spin = 0; !mbar_try_wait(bar, parity); ++spin)
    if (spin > (1u << 24)) __trap();
}
__device__ __forceinline__ void tma_load_5d(unsigned dst, const CUtensorMap* map, unsigned bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void cp_async16_zfill(unsigned dst, const void* src, bool pred) {
  const int n = pred ? 16 : 0;  // src-size 0: the 16 destination bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ uint4 lds128(unsigned addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}

// ---------------------------------------------------------------------------------------------------------------
// Sample position: pixel-space coordinate of one axis, M2F:807 + grid_sampler_unnormalize(align_corners=False),
// evaluated in the reference's order without FMA contraction (same as axis_setup in msda_common.cuh).
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float pixel_coord(float coord, int n) {
  const float g = __fadd_rn(__fmul_rn(2.f, coord), -1.f);
  return __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(g, 1.f), (float)n), -1.f), 0.5f);
}

struct SamplePos {
  int ix0, iy0;   // top-left pixel of the footprint, in [-1, W-1] x [-1, H-1] when live
  float fx, fy;   // fractional parts
  bool live;      // at least one corner can lie inside the level (false for NaN / far away)
};
__device__ __forceinline__ SamplePos sample_pos(float2 xy, int W, int H) {
  SamplePos s;
  const float px = pixel_coord(xy.x, W), py = pixel_coord(xy.y, H);
  s.live = (px >= -1.f) && (px < (float)W) && (py >= -1.f) && (py < (float)H);  // false for NaN
  const float flx = floorf(px), fly = floorf(py);
  s.fx = px - flx;
  s.fy = py - fly;
  s.ix0 = s.live ? (int)flx : -1;
  s.iy0 = s.live ? (int)fly : -1;
  return s;
}

// Window origin along one axis: the bounding box [mn, mx + 1] of the footprints if it fits into `box` pixels,
// otherwise the box is centred inside it (samples outside take the slow path).
__device__ __forceinline__ int window_origin(int mn, int mx, int box) {
  const int extent = mx + 2 - mn;
  return extent <= box ? mn : mn + (extent - box) / 2;
}


struct SmemLayout {
  int guard, desc, qidx, bars, bbox, total;
};
// guard: all-zero bytes a window-less sample reads (top row at 0, bottom row at + pitch): max pitch + 128 bytes
__host__ __device__ inline SmemLayout win_smem_layout(int win_bytes, int guard_bytes) {
  SmemLayout s;
  int o = (win_bytes + 127) & ~127;
  s.guard = o; o += (guard_bytes + 127) & ~127;
  s.desc = o; o += 2 * kWinP * kWinTQ * 16;
  s.qidx = o; o += kWinTQ * 4;
  s.bars = o; o += kWinL * 8;
  s.bbox = o; o += kWinL * 4 * 4;
  s.total = o;
  return s;
}

// optional phase timing (dev tool): thread 0 of the first kWinDbgBlocks blocks records clock64() after each phase
constexpr int kWinDbgBlocks = 512, kWinDbgSlots = 12;
__device__ __forceinline__ void dbg_mark(long long* dbg, int slot) {
  if (dbg && threadIdx.x == 0 && blockIdx.x < kWinDbgBlocks) dbg[blockIdx.x * kWinDbgSlots + slot] = clock64();
}

// ---------------------------------------------------------------------------------------------------------------
// Forward
// ---------------------------------------------------------------------------------------------------------------
// STAGE 0: TMA box loads; STAGE 1: per-thread cp.async (LDGSTS) with zero fill -- same shared-memory image.
template <typename AT, int L, int STAGE>
__global__ void __launch_bounds__(kWinNT, 2) msda_fwd_win_kernel(const __grid_constant__ WParams p,
                                                                 const __grid_constant__ WinMaps maps) {
  constexpr int NT = kWinNT, P = kWinP, TQ = kWinTQ;
  constexpr int SPT = (TQ * P + NT - 1) / NT;  // samples per thread and level
  constexpr int NG = NT / 8, QPG = TQ / NG;    // lane groups, queries per lane group
  extern __shared__ __align__(128) unsigned char smem[];
  const SmemLayout lay = win_smem_layout(p.win_bytes, p.guard_bytes);
  uint4* s_desc = reinterpret_cast<uint4*>(smem + lay.desc);  // [2][P][TQ]
  int* s_q = reinterpret_cast<int*>(smem + lay.qidx);         // [TQ] query index, -1 beyond the tile
  int* s_bbox = reinterpret_cast<int*>(smem + lay.bbox);      // [L][4] = minx, maxx, miny, maxy
  const unsigned bar0 = smem_u32(smem + lay.bars);
  const unsigned win0 = smem_u32(smem);

  const KParams& k = p.k;
  int bid = blockIdx.x;
  const int h = bid % k.H;
  bid /= k.H;
  const int tile = bid % k.num_tiles;
  const int b = bid / k.num_tiles;
  const int q0 = p.tile_start[tile];
  const int nq = min(p.tile_start[tile + 1] - q0, TQ);
  const int tid = threadIdx.x, lane = tid & 31;
  dbg_mark(p.dbg, 0);

  if (tid < TQ) s_q[tid] = tid < nq ? k.q_order[q0 + tid] : -1;
  if (tid < L * 4) s_bbox[tid] = (tid & 1) ? -0x7fffffff : 0x7fffffff;
  for (int i = tid; i < p.guard_bytes / 16; i += NT) reinterpret_cast<uint4*>(smem + lay.guard)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (tid == 0 && STAGE == 0) {
#pragma unroll
    for (int l = 0; l < L; ++l) mbar_init(bar0 + 8 * l, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // ---- pass 1a: every location / weight of this thread's samples, all loads in flight together
  float2 lc[L][SPT];
  float av[L][SPT];
#pragma unroll
  for (int r = 0; r < SPT; ++r) {
    const int i = r * NT + tid;
    const int q = s_q[i / P];
    const long long base = (((long long)b * k.Q + max(q, 0)) * k.H + h) * k.LP + (i % P);
#pragma unroll
    for (int l = 0; l < L; ++l) {
      lc[l][r] = make_float2(__int_as_float(0x7fc00000), 0.f);  // NaN: not live
      av[l][r] = 0.f;
      if (q >= 0) {
        lc[l][r] = __ldg(reinterpret_cast<const float2*>(k.loc) + base + l * P);
        av[l][r] = to_float<AT>(reinterpret_cast<const AT*>(k.attn)[base + l * P]);
      }
    }
  }
  // ---- pass 1b: sample positions (kept in registers) and the per-level bounding boxes
  int pk[L][SPT];
  float s_fx[L][SPT], s_wt[L][SPT], s_wb[L][SPT];
#pragma unroll
  for (int l = 0; l < L; ++l) {
    const int W = k.lv[l].W, H = k.lv[l].H;
    int mnx = 0x7fffffff, mxx = -0x7fffffff, mny = 0x7fffffff, mxy = -0x7fffffff;
#pragma unroll
    for (int r = 0; r < SPT; ++r) {
      const SamplePos sp = sample_pos(lc[l][r], W, H);
      const float a = av[l][r];
      s_fx[l][r] = sp.fx;
      s_wt[l][r] = a * (1.f - sp.fy);
      s_wb[l][r] = a * sp.fy;
      pk[l][r] = (sp.ix0 + 1) | ((sp.iy0 + 1) << 13) | (sp.live ? 0 : (int)kFlagDead);
      if (sp.live) {
        mnx = min(mnx, sp.ix0); mxx = max(mxx, sp.ix0);
        mny = min(mny, sp.iy0); mxy = max(mxy, sp.iy0);
      }
    }
    mnx = __reduce_min_sync(0xffffffffu, mnx); mxx = __reduce_max_sync(0xffffffffu, mxx);
    mny = __reduce_min_sync(0xffffffffu, mny); mxy = __reduce_max_sync(0xffffffffu, mxy);
    if (lane == 0 && mxx >= mnx) {
      atomicMin(&s_bbox[l * 4 + 0], mnx); atomicMax(&s_bbox[l * 4 + 1], mxx);
      atomicMin(&s_bbox[l * 4 + 2], mny); atomicMax(&s_bbox[l * 4 + 3], mxy);
    }
  }
  __syncthreads();
  dbg_mark(p.dbg, 1);

  // ---- windows: origin per level (same value in every thread), then the asynchronous box loads
  int wx0[L], wy0[L];
  bool has_win[L];
#pragma unroll
  for (int l = 0; l < L; ++l) {
    const int mnx = s_bbox[l * 4 + 0], mxx = s_bbox[l * 4 + 1], mny = s_bbox[l * 4 + 2], mxy = s_bbox[l * 4 + 3];
    has_win[l] = mxx >= mnx;
    wx0[l] = has_win[l] ? window_origin(mnx, mxx, p.bw[l]) : 0;
    wy0[l] = has_win[l] ? window_origin(mny, mxy, p.bh[l]) : 0;
  }
  if (STAGE == 0) {
    if (tid == 0) {
#pragma unroll
      for (int l = 0; l < L; ++l)
        if (has_win[l]) {
          mbar_expect_tx(bar0 + 8 * l, (unsigned)(p.bw[l] * p.bh[l] * kRowB));
          tma_load_5d(win0 + p.woff[l], &maps.m[l], bar0 + 8 * l, 0, h, wx0[l], wy0[l], b);
        }
    }
  } else {
#pragma unroll
    for (int l = 0; l < L; ++l) {
      if (has_win[l]) {
        const int W = k.lv[l].W, H = k.lv[l].H, bw = p.bw[l];
        const int n16 = bw * p.bh[l] * (kRowB / 16);
        const uint4* src0 = reinterpret_cast<const uint4*>(k.value) + (long long)b * k.batch_stride16;
        for (int i = tid; i < n16; i += NT) {
          const int pix = i >> 2, c = i & 3;
          const int wy = pix / bw, wx = pix - wy * bw;
          const int gx = wx0[l] + wx, gy = wy0[l] + wy;
          const bool in = (unsigned)gx < (unsigned)W && (unsigned)gy < (unsigned)H;
          const long long u = in ? ((long long)(k.lv[l].start + gy * W + gx) * k.H + h) * 4 + c : 0;
          cp_async16_zfill(win0 + p.woff[l] + i * 16, src0 + u, in);
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
  }

  // ---- pass 2 helper: descriptors of level l -> buffer (l & 1): {window offset, a*(1-fy), a*fy, fx}. A sample whose
  // footprint is not inside the window keeps its weights but points at the all-zero guard rows (so the gather loop
  // needs no branch) and carries its level coordinates in the sign-flagged first word for the slow pass.
  auto write_desc = [&](int l) {
    uint4* dst = s_desc + (l & 1) * (P * TQ);
    const int bw = p.bw[l], bh = p.bh[l];
#pragma unroll
    for (int r = 0; r < SPT; ++r) {
      const int i = r * NT + tid;
      const int ql = i / P, pt = i % P;
      const int code = pk[l][r];
      const int ix0 = (code & 0x1fff) - 1, iy0 = ((code >> 13) & 0x1fff) - 1;
      const int ox = ix0 - wx0[l], oy = iy0 - wy0[l];
      const bool inwin = has_win[l] && !(code & (int)kFlagDead) && (unsigned)ox <= (unsigned)(bw - 2) &&
                         (unsigned)oy <= (unsigned)(bh - 2);
      const unsigned first = inwin ? (unsigned)((oy * bw + ox) * kRowB) : ((unsigned)code | kFlagSlow);
      dst[pt * TQ + ql] = make_uint4(first, __float_as_uint(s_wt[l][r]), __float_as_uint(s_wb[l][r]),
                                     __float_as_uint(s_fx[l][r]));
    }
  };

  // ---- gather: 8 lanes per (query, head); lanes 0-3 take the left pixel of a pair, 4-7 the right one, 16 B each
  const int g = tid >> 3, l8 = tid & 7, side = l8 >> 2, chunk = l8 & 3;
  const float sgn = side ? 1.f : -1.f, bas = side ? 0.f : 1.f;  // wx = side ? fx : 1 - fx
  float acc[QPG][8];
#pragma unroll
  for (int qi = 0; qi < QPG; ++qi)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[qi][j] = 0.f;

  write_desc(0);
  __syncthreads();
  dbg_mark(p.dbg, 2);
#pragma unroll
  for (int l = 0; l < L; ++l) {
    if (l + 1 < L) write_desc(l + 1);
    if (has_win[l]) {
      if (STAGE == 0) {
        mbar_wait(bar0 + 8 * l, 0);
      } else {
        // groups are committed in level order; level l is complete once at most L-1-l newer groups are pending
        if (l == 0) asm volatile("cp.async.wait_group %0;" ::"n"(L - 1) : "memory");
        if (l == 1) asm volatile("cp.async.wait_group %0;" ::"n"(L > 2 ? L - 2 : 0) : "memory");
        if (l == 2) asm volatile("cp.async.wait_group %0;" ::"n"(L > 3 ? L - 3 : 0) : "memory");
        if (l >= 3) asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();  // every thread's copies have landed
      }
    }
    dbg_mark(p.dbg, 3 + 2 * l);
    const uint4* dsc = s_desc + (l & 1) * (P * TQ);
    const unsigned wbase = win0 + p.woff[l] + l8 * 16;
    const unsigned guard = (unsigned)(lay.guard - p.woff[l]);  // guard rows relative to this level's window
    const unsigned pitch = p.bw[l] * kRowB;
    const Level lv = k.lv[l];
#pragma unroll
    for (int qi = 0; qi < QPG; ++qi) {
      const int ql = g + qi * NG;
      if (ql < nq) {
        uint4 d[P];
#pragma unroll
        for (int pt = 0; pt < P; ++pt) d[pt] = dsc[pt * TQ + ql];
        unsigned any = 0;
#pragma unroll
        for (int pt = 0; pt < P; ++pt) {
          any |= d[pt].x;
          const unsigned off = (int)d[pt].x >= 0 ? d[pt].x : guard;
          const uint4 top = lds128(wbase + off), bot = lds128(wbase + off + pitch);
          const float wx = fmaf(sgn, __uint_as_float(d[pt].w), bas);
          const float wt = __uint_as_float(d[pt].y) * wx, wb = __uint_as_float(d[pt].z) * wx;
          float f[8];
          Vec16<__nv_bfloat16>::unpack(top, f);
#pragma unroll
          for (int j = 0; j < 8; j += 2) fma2_scalar(acc[qi][j], acc[qi][j + 1], wt, f[j], f[j + 1]);
          Vec16<__nv_bfloat16>::unpack(bot, f);
#pragma unroll
          for (int j = 0; j < 8; j += 2) fma2_scalar(acc[qi][j], acc[qi][j + 1], wb, f[j], f[j + 1]);
        }
        if ((int)any < 0) {
          // slow pass: footprints outside the window, bounds-checked loads from global memory
#pragma unroll
          for (int pt = 0; pt < P; ++pt) {
            const unsigned c = d[pt].x;
            if ((int)c < 0 && !(c & kFlagDead)) {
              const int x = (int)(c & 0x1fff) - 1 + side, y0 = (int)((c >> 13) & 0x1fff) - 1;
              if ((unsigned)x < (unsigned)lv.W) {
                const uint4* vb = reinterpret_cast<const uint4*>(k.value) + (long long)b * k.batch_stride16 + chunk;
                const float wx = fmaf(sgn, __uint_as_float(d[pt].w), bas);
                float f[8];
                if ((unsigned)y0 < (unsigned)lv.H) {
                  const float wt = __uint_as_float(d[pt].y) * wx;
                  Vec16<__nv_bfloat16>::unpack(ldg16(vb + ((long long)(lv.start + y0 * lv.W + x) * k.H + h) * 4), f);
#pragma unroll
                  for (int j = 0; j < 8; j += 2) fma2_scalar(acc[qi][j], acc[qi][j + 1], wt, f[j], f[j + 1]);
                }
                if ((unsigned)(y0 + 1) < (unsigned)lv.H) {
                  const float wb = __uint_as_float(d[pt].z) * wx;
                  Vec16<__nv_bfloat16>::unpack(ldg16(vb + ((long long)(lv.start + (y0 + 1) * lv.W + x) * k.H + h) * 4), f);
#pragma unroll
                  for (int j = 0; j < 8; j += 2) fma2_scalar(acc[qi][j], acc[qi][j + 1], wb, f[j], f[j + 1]);
                }
              }
            }
          }
        }
      }
    }
    dbg_mark(p.dbg, 4 + 2 * l);
    __syncthreads();  // descriptors of level l+1 visible; buffer (l & 1) free for level l+2
  }

  // ---- left + right halves, one 16-byte store per lane of the left half
#pragma unroll
  for (int qi = 0; qi < QPG; ++qi) {
    const int ql = g + qi * NG;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[qi][j] += __shfl_xor_sync(0xffffffffu, acc[qi][j], 4);
    if (ql < nq && side == 0) {
      const int q = s_q[ql];
      uint4* o = reinterpret_cast<uint4*>(k.out) + (((long long)b * k.Q + q) * k.H + h) * 4 + chunk;
      *o = Vec16<__nv_bfloat16>::pack(acc[qi]);
    }
  }
  dbg_mark(p.dbg, 11);
}

// ---------------------------------------------------------------------------------------------------------------
// Forward on the tensor cores (mma.sync m16n8k16, bf16 x bf16 -> fp32)
// ---------------------------------------------------------------------------------------------------------------
// One warp owns one (query, head) at a time. Per level the 16 bilinear corners of the query's four samples are the K
// dimension of one MMA pair:  D[channel, col] += sum_corner V[corner][channel] * Wt[corner][col].
//   A (16 channels x 16 corners) comes straight out of the staged window with ldmatrix.x4.trans: every lane supplies
//     the address of one 16-byte row (8 channels of one corner), so the gather needs no unpacking and no FMAs.
//   B (16 corners x 8 columns) holds the bilinear weight * attention weight of the corners, pre-packed per sample by
//     the descriptor pass as bf16 hi and bf16 lo = w - hi (columns 0-3 / 4-7: the products are exact in fp32, so
//     hi + lo reproduces the fp32 weight to 2^-17).
// Bank conflicts: the eight rows of one 8x8 matrix are eight different corners reading the SAME 16-byte chunk index,
// and corner rows start at 0 or 64 mod 128 -- a 4-way conflict. So sample p reads channel chunk (v + p) & 3 where the
// fragment asks for chunk v: the matrix rows then cover all eight 16-byte slots of a 128-byte bank line exactly once
// (two sides x four samples). The product is only meaningful where the weight column uses the same rotation, hence
// column c of B is non-zero only for the corners of sample c & 3, and D[v, c] belongs to real chunk (v + c) & 3;
// the epilogue adds the columns up and leaves lane (g, t) with output channel t * 8 + g.
// A sample whose footprint is not in the window reads the zero guard rows here and is added by a scalar pass.
struct MmaLayout {
  int guard, w, off, qidx, bars, bbox, total;
};
__host__ __device__ inline MmaLayout mma_smem_layout(int win_bytes, int guard_bytes, int L, int tqs) {
  MmaLayout s;
  int o = (win_bytes + 127) & ~127;
  s.guard = o; o += (guard_bytes + 127) & ~127;
  s.w = o;     o += L * kWinP * tqs * 16;
  s.off = o;   o += (L * kWinP * tqs * 2 + 15) & ~15;
  s.qidx = o;  o += (tqs * 4 + 15) & ~15;
  s.bars = o;  o += kWinL * 8;
  s.bbox = o;  o += kWinL * 4 * 4;
  s.total = o;
  return s;
}

__device__ __forceinline__ void ldsm_x4_trans(unsigned addr, unsigned (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ unsigned pack_bf16x2(float lo, float hi) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const unsigned*>(&h);
}

template <typename AT, int L, int NT>
__global__ void __launch_bounds__(NT, 2) msda_fwd_mma_kernel(const __grid_constant__ WParams p,
                                                             const __grid_constant__ WinMaps maps) {
  constexpr int P = kWinP, NW = NT / 32;
  constexpr int SPT = (kWinTQ * P + NT - 1) / NT;  // samples per thread and level
  constexpr int QU = 2;                            // queries in flight per warp
  extern __shared__ __align__(128) unsigned char smem[];
  const int TQS = p.tqs;
  const MmaLayout lay = mma_smem_layout(p.win_bytes, p.guard_bytes, L, TQS);
  uint4* s_w = reinterpret_cast<uint4*>(smem + lay.w);                      // [L*P][TQS] weight packs
  unsigned short* s_off = reinterpret_cast<unsigned short*>(smem + lay.off);  // [L*P][TQS] window offset / 64, 0xffff = slow
  int* s_q = reinterpret_cast<int*>(smem + lay.qidx);
  int* s_bbox = reinterpret_cast<int*>(smem + lay.bbox);
  const unsigned bar0 = smem_u32(smem + lay.bars);
  const unsigned win0 = smem_u32(smem);

  // block = (batch element, tile, group of `hpb` heads): the tile's query list is read once, and the locations of
  // head h + 1 are fetched into registers while head h is gathered
  const KParams& k = p.k;
  const int hgroups = k.H / p.hpb;
  int bid = blockIdx.x;
  const int h_begin = (bid % hgroups) * p.hpb;
  bid /= hgroups;
  const int tile = bid % k.num_tiles;
  const int b = bid / k.num_tiles;
  const int q0 = p.tile_start[tile];
  const int nq = min(p.tile_start[tile + 1] - q0, TQS);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  dbg_mark(p.dbg, 0);

  for (int i = tid; i < TQS; i += NT) s_q[i] = i < nq ? k.q_order[q0 + i] : -1;
  for (int i = tid; i < p.guard_bytes / 16; i += NT) reinterpret_cast<uint4*>(smem + lay.guard)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (tid == 0) {
#pragma unroll
    for (int l = 0; l < L; ++l) mbar_init(bar0 + 8 * l, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // ---- lane constants of the fragment loads
  const int g = lane >> 2, t = lane & 3;
  const bool b_active = (g & 3) == t;                 // this lane holds the B entries of sample t (hi: g < 4, lo: g >= 4)
  const int mj = lane >> 3, mr = lane & 7;            // ldmatrix: matrix index, row inside the matrix
  const int a_p = mr >> 1, a_s = mr & 1, a_row = mj >> 1;  // sample, side, top/bottom of the corner this lane addresses
  unsigned bw_base[L], bo_base[L], a_base0[L], a_base1[L], guard_off[L];
#pragma unroll
  for (int l = 0; l < L; ++l) {
    bw_base[l] = b_active ? smem_u32(s_w + (l * P + t) * TQS) + ((g >> 2) << 3) : win0 + lay.guard;
    bo_base[l] = smem_u32(s_off + (l * P + a_p) * TQS);
    const unsigned c = win0 + p.woff[l] + a_row * (p.bw[l] * kRowB) + a_s * 64;
    a_base0[l] = c + ((((mj & 1) + a_p) & 3) << 4);
    a_base1[l] = c + ((((mj & 1) + 2 + a_p) & 3) << 4);
    guard_off[l] = (unsigned)(lay.guard - p.woff[l]);
  }
  const unsigned b_stride = b_active ? 16u : 0u;

  // ---- this thread's samples: i = r * NT + tid -> (query i / P, point i % P), the same for every head
  long long sbase[SPT];
#pragma unroll
  for (int r = 0; r < SPT; ++r) {
    const int i = r * NT + tid;
    const int ql = i / P;
    const int q = ql < TQS ? s_q[ql] : -1;
    sbase[r] = q >= 0 ? (((long long)b * k.Q + q) * k.H) * k.LP + (i % P) : -1;
  }
  float2 lc[L][SPT];
  float av[L][SPT];
  auto fetch = [&](int h) {
#pragma unroll
    for (int r = 0; r < SPT; ++r)
#pragma unroll
      for (int l = 0; l < L; ++l) {
        lc[l][r] = make_float2(__int_as_float(0x7fc00000), 0.f);  // NaN: not live
        av[l][r] = 0.f;
        if (sbase[r] >= 0) {
          const long long si = sbase[r] + (long long)h * k.LP + l * P;
          lc[l][r] = __ldg(reinterpret_cast<const float2*>(k.loc) + si);
          av[l][r] = to_float<AT>(reinterpret_cast<const AT*>(k.attn)[si]);
        }
      }
  };
  fetch(h_begin);
  unsigned phase[L];  // mbarrier phase per level (a level without live samples skips its load)
#pragma unroll
  for (int l = 0; l < L; ++l) phase[l] = 0;

  for (int hi = 0; hi < p.hpb; ++hi) {
    const int h = h_begin + hi;
    if (tid < L * 4) s_bbox[tid] = (tid & 1) ? -0x7fffffff : 0x7fffffff;
    __syncthreads();  // bbox reset visible; previous head's gather finished with s_w / s_off / windows

    // ---- sample positions and the per-level bounding boxes
    int pk[L][SPT];
    float s_fx[L][SPT], s_wt[L][SPT], s_wb[L][SPT];
#pragma unroll
    for (int l = 0; l < L; ++l) {
      const int W = k.lv[l].W, H = k.lv[l].H;
      int mnx = 0x7fffffff, mxx = -0x7fffffff, mny = 0x7fffffff, mxy = -0x7fffffff;
#pragma unroll
      for (int r = 0; r < SPT; ++r) {
        const SamplePos sp = sample_pos(lc[l][r], W, H);
        const float a = av[l][r];
        s_fx[l][r] = sp.fx;
        s_wt[l][r] = a * (1.f - sp.fy);
        s_wb[l][r] = a * sp.fy;
        pk[l][r] = (sp.ix0 + 1) | ((sp.iy0 + 1) << 13) | (sp.live ? 0 : (int)kFlagDead);
        if (sp.live) {
          mnx = min(mnx, sp.ix0); mxx = max(mxx, sp.ix0);
          mny = min(mny, sp.iy0); mxy = max(mxy, sp.iy0);
        }
      }
      mnx = __reduce_min_sync(0xffffffffu, mnx); mxx = __reduce_max_sync(0xffffffffu, mxx);
      mny = __reduce_min_sync(0xffffffffu, mny); mxy = __reduce_max_sync(0xffffffffu, mxy);
      if (lane == 0 && mxx >= mnx) {
        atomicMin(&s_bbox[l * 4 + 0], mnx); atomicMax(&s_bbox[l * 4 + 1], mxx);
        atomicMin(&s_bbox[l * 4 + 2], mny); atomicMax(&s_bbox[l * 4 + 3], mxy);
      }
    }
    __syncthreads();
    if (hi == 0) dbg_mark(p.dbg, 1);

    // ---- windows: origin per level, TMA box loads (zero fill outside the level)
    int wx0[L], wy0[L];
    bool has_win[L];
#pragma unroll
    for (int l = 0; l < L; ++l) {
      const int mnx = s_bbox[l * 4 + 0], mxx = s_bbox[l * 4 + 1], mny = s_bbox[l * 4 + 2], mxy = s_bbox[l * 4 + 3];
      has_win[l] = mxx >= mnx;
      wx0[l] = has_win[l] ? window_origin(mnx, mxx, p.bw[l]) : 0;
      wy0[l] = has_win[l] ? window_origin(mny, mxy, p.bh[l]) : 0;
    }
    if (tid == 0) {
#pragma unroll
      for (int l = 0; l < L; ++l)
        if (has_win[l]) {
          mbar_expect_tx(bar0 + 8 * l, (unsigned)(p.bw[l] * p.bh[l] * kRowB));
          tma_load_5d(win0 + p.woff[l], &maps.m[l], bar0 + 8 * l, 0, h, wx0[l], wy0[l], b);
        }
    }

    // ---- descriptors: weight packs {hi(top pair), hi(bottom pair), lo(top pair), lo(bottom pair)}, window offsets
#pragma unroll
    for (int l = 0; l < L; ++l) {
      const int bw = p.bw[l], bh = p.bh[l];
#pragma unroll
      for (int r = 0; r < SPT; ++r) {
        const int i = r * NT + tid;
        const int ql = i / P, pt = i % P;
        if (ql < TQS) {
          const int code = pk[l][r];
          const int ix0 = (code & 0x1fff) - 1, iy0 = ((code >> 13) & 0x1fff) - 1;
          const int ox = ix0 - wx0[l], oy = iy0 - wy0[l];
          const bool inwin = has_win[l] && !(code & (int)kFlagDead) && (unsigned)ox <= (unsigned)(bw - 2) &&
                             (unsigned)oy <= (unsigned)(bh - 2);
          const int slot = (l * P + pt) * TQS + ql;
          uint4 wv = make_uint4(0u, 0u, 0u, 0u);  // not in the window: the MMA sees zero weights, the scalar pass redoes it
          if (inwin) {
            const float fx = s_fx[l][r], gx = 1.f - fx;
            const float w00 = s_wt[l][r] * gx, w01 = s_wt[l][r] * fx, w10 = s_wb[l][r] * gx, w11 = s_wb[l][r] * fx;
            wv.x = pack_bf16x2(w00, w01);
            wv.y = pack_bf16x2(w10, w11);
            wv.z = pack_bf16x2(w00 - __uint_as_float(wv.x << 16), w01 - __uint_as_float(wv.x & 0xffff0000u));
            wv.w = pack_bf16x2(w10 - __uint_as_float(wv.y << 16), w11 - __uint_as_float(wv.y & 0xffff0000u));
          }
          s_w[slot] = wv;
          s_off[slot] = inwin ? (unsigned short)(oy * bw + ox) : (unsigned short)0xffffu;
        }
      }
    }
    __syncthreads();
    if (hi == 0) dbg_mark(p.dbg, 2);
    if (hi + 1 < p.hpb) fetch(h + 1);  // in flight during the gather
#pragma unroll
    for (int l = 0; l < L; ++l)
      if (has_win[l]) {
        mbar_wait(bar0 + 8 * l, phase[l]);
        phase[l] ^= 1;
      }
    if (hi == 0) dbg_mark(p.dbg, 3);

    // ---- gather: QU queries per warp step, one accumulator pair per level (no dependent MMA chain)
    for (int qb = warp; qb < nq; qb += NW * QU) {
      float d0[QU][L][4], d1[QU][L][4];
      unsigned slow[QU];
      int qls[QU];
#pragma unroll
      for (int u = 0; u < QU; ++u) {
        qls[u] = min(qb + u * NW, nq - 1);  // a clamped duplicate in the ragged tail; its result is not stored
        slow[u] = 0;
#pragma unroll
        for (int l = 0; l < L; ++l) {
          unsigned b0, b1;
          asm("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(b0), "=r"(b1) : "r"(bw_base[l] + qls[u] * b_stride));
          unsigned short o16;
          asm("ld.shared.u16 %0, [%1];" : "=h"(o16) : "r"(bo_base[l] + qls[u] * 2));
          const bool is_slow = o16 == 0xffffu;
          slow[u] |= is_slow ? (1u << l) : 0u;
          const unsigned off = is_slow ? guard_off[l] : (unsigned)o16 * kRowB;
          unsigned a0[4], a1[4];
          ldsm_x4_trans(a_base0[l] + off, a0);
          ldsm_x4_trans(a_base1[l] + off, a1);
#pragma unroll
          for (int j = 0; j < 4; ++j) d0[u][l][j] = d1[u][l][j] = 0.f;
          mma_bf16_16816(d0[u][l], a0, b0, b1);
          mma_bf16_16816(d1[u][l], a1, b0, b1);
        }
      }
#pragma unroll
      for (int u = 0; u < QU; ++u) {
        // epilogue: columns -> channel chunks. E[j] sums the two columns of this lane over the levels; its real chunk
        // is j (t even) or j ^ 2 (t odd); a reduce-scatter over the four lanes of a row leaves chunk t in lane (g, t)
        float e0 = 0.f, e1 = 0.f, e2 = 0.f, e3 = 0.f;
#pragma unroll
        for (int l = 0; l < L; ++l) {
          e0 += d0[u][l][0] + d1[u][l][3];
          e1 += d0[u][l][2] + d0[u][l][1];
          e2 += d1[u][l][0] + d0[u][l][3];
          e3 += d1[u][l][2] + d1[u][l][1];
        }
        const bool odd = t & 1, up = t & 2;
        const float c0 = odd ? e2 : e0, c1 = odd ? e3 : e1, c2 = odd ? e0 : e2, c3 = odd ? e1 : e3;
        const float k0 = (up ? c2 : c0) + __shfl_xor_sync(0xffffffffu, up ? c0 : c2, 2);
        const float k1 = (up ? c3 : c1) + __shfl_xor_sync(0xffffffffu, up ? c1 : c3, 2);
        float o = (odd ? k1 : k0) + __shfl_xor_sync(0xffffffffu, odd ? k0 : k1, 1);
        const int ql = qls[u];
        const int q = s_q[ql];
        const int ch = t * 8 + g;
        const unsigned sl = __reduce_or_sync(0xffffffffu, slow[u]);
        if (sl) {
          // scalar pass for the samples outside the window (location and weight are read again; the whole warp works
          // on one sample, this lane on channel `ch`): bounds-checked corner loads from global memory
          const __nv_bfloat16* vb = reinterpret_cast<const __nv_bfloat16*>(k.value) + (long long)b * k.batch_stride16 * 8 + h * 32 + ch;
          const long long sb = (((long long)b * k.Q + q) * k.H + h) * k.LP;
#pragma unroll
          for (int l = 0; l < L; ++l) {
            if (!((sl >> l) & 1)) continue;
            const Level lv = k.lv[l];
            for (int pt = 0; pt < P; ++pt) {
              if (s_off[(l * P + pt) * TQS + ql] != 0xffffu) continue;
              const SamplePos sp = sample_pos(__ldg(reinterpret_cast<const float2*>(k.loc) + sb + l * P + pt), lv.W, lv.H);
              if (!sp.live) continue;
              const float a = to_float<AT>(reinterpret_cast<const AT*>(k.attn)[sb + l * P + pt]);
              const float wt = a * (1.f - sp.fy), wb = a * sp.fy;
#pragma unroll
              for (int cn = 0; cn < 4; ++cn) {
                const int x = sp.ix0 + (cn & 1), y = sp.iy0 + (cn >> 1);
                if ((unsigned)x < (unsigned)lv.W && (unsigned)y < (unsigned)lv.H) {
                  const float w = ((cn >> 1) ? wb : wt) * ((cn & 1) ? sp.fx : 1.f - sp.fx);
                  o = fmaf(w, __bfloat162float(vb[(long long)(lv.start + y * lv.W + x) * (k.H * 32)]), o);
                }
              }
            }
          }
        }
        if (u == 0 || qb + u * NW < nq)
          reinterpret_cast<__nv_bfloat16*>(k.out)[(((long long)b * k.Q + q) * k.H + h) * 32 + ch] = __float2bfloat16_rn(o);
      }
    }
    if (hi == 0) dbg_mark(p.dbg, 11);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      f = nullptr;
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  return fn;
}

int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v && *v ? atoi(v) : dflt;
}

// Window boxes: the tile's patch scaled to the level plus `halo` pixels on every side (covers sampling offsets up to
// halo - 0.5 px together with the +1 of the bilinear footprint); never larger than the level plus its one-pixel zero
// border. Shrinks the halo until the windows fit.
bool plan_windows(const msda_b200_desc* d, WParams& w, int budget_bytes, bool mma) {
  const int L = d->L;
  int fine = 0;
  for (int l = 1; l < L; ++l)
    if ((long long)d->spatial_shapes_hw[2 * l] * d->spatial_shapes_hw[2 * l + 1] >
        (long long)d->spatial_shapes_hw[2 * fine] * d->spatial_shapes_hw[2 * fine + 1])
      fine = l;
  const int Hf = d->spatial_shapes_hw[2 * fine], Wf = d->spatial_shapes_hw[2 * fine + 1];
  const int th = d->tile_rows > 0 ? d->tile_rows : 8, tw = d->tile_cols > 0 ? d->tile_cols : 16;
  for (int halo = env_int("MSDA_B200_WIN_HALO", 6); halo >= 1; --halo) {
    int off = 0;
    for (int l = 0; l < L; ++l) {
      const int Hl = d->spatial_shapes_hw[2 * l], Wl = d->spatial_shapes_hw[2 * l + 1];
      const int ph = (int)(((long long)th * Hl + Hf - 1) / Hf), pw = (int)(((long long)tw * Wl + Wf - 1) / Wf);
      int bh = ph + 2 * halo, bw = pw + 2 * halo;
      if (bh > Hl + 2) bh = Hl + 2;
      if (bw > Wl + 2) bw = Wl + 2;
      if (bh < 2) bh = 2;
      if (bw < 2) bw = 2;
      if (bh > 256 || bw > 256) return false;
      w.bw[l] = bw; w.bh[l] = bh; w.woff[l] = off;
      off += (bw * bh * kRowB + 127) & ~127;
    }
    w.win_bytes = off;
    int maxbw = 0;
    for (int l = 0; l < L; ++l) maxbw = w.bw[l] > maxbw ? w.bw[l] : maxbw;
    w.guard_bytes = maxbw * kRowB + 128;
    const int need = mma ? mma_smem_layout(off, w.guard_bytes, L, w.tqs).total : win_smem_layout(off, w.guard_bytes).total;
    if (need <= budget_bytes) return true;
  }
  return false;
}

bool encode_maps(const msda_b200_desc* d, const void* value, const WParams& w, WinMaps& maps) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return false;
  for (int l = 0; l < d->L; ++l) {
    const cuuint64_t Hl = (cuuint64_t)d->spatial_shapes_hw[2 * l], Wl = (cuuint64_t)d->spatial_shapes_hw[2 * l + 1];
    const cuuint64_t row = (cuuint64_t)d->H * kRowB;  // bytes of one pixel (all heads)
    const cuuint64_t gdim[5] = {32, (cuuint64_t)d->H, Wl, Hl, (cuuint64_t)d->B};
    const cuuint64_t gstr[4] = {kRowB, row, Wl * row, (cuuint64_t)d->S * row};
    const cuuint32_t box[5] = {32, 1, (cuuint32_t)w.bw[l], (cuuint32_t)w.bh[l], 1};
    const cuuint32_t est[5] = {1, 1, 1, 1, 1};
    char* base = const_cast<char*>(reinterpret_cast<const char*>(value)) + (size_t)d->level_start_index[l] * row;
    CUresult r = enc(&maps.m[l], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, base, gdim, gstr, box, est,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)  // the sibling heads' blocks want the neighbouring bytes anyway, but promotion is optional
      r = enc(&maps.m[l], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, base, gdim, gstr, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return false;
  }
  return true;
}

template <typename AT, int STAGE>
int launch_fwd_win(const msda_b200_desc* d, const WParams& w, const WinMaps& maps, cudaStream_t st) {
  const size_t smem = win_smem_layout(w.win_bytes, w.guard_bytes).total;
  const long long blocks = (long long)w.k.B * w.k.num_tiles * w.k.H;
  if (blocks > 0x7fffffffll) return msda_b200_internal_fail(MSDA_B200_ERR_UNSUPPORTED, "forward (window): grid too large");
#define MSDA_WIN_LAUNCH(LV)                                                                                           \
  {                                                                                                                   \
    auto kern = msda_fwd_win_kernel<AT, LV, STAGE>;                                                                   \
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)            \
      return msda_b200_internal_fail(MSDA_B200_ERR_CUDA, "forward (window): cannot reserve shared memory");           \
    kern<<<(unsigned)blocks, kWinNT, smem, st>>>(w, maps);                                                            \
  }
  switch (d->L) {
    case 1: MSDA_WIN_LAUNCH(1) break;
    case 2: MSDA_WIN_LAUNCH(2) break;
    case 3: MSDA_WIN_LAUNCH(3) break;
    case 4: MSDA_WIN_LAUNCH(4) break;
    default: return msda_b200_internal_fail(MSDA_B200_ERR_UNSUPPORTED, "forward (window): L > 4");
  }
#undef MSDA_WIN_LAUNCH
  return MSDA_B200_OK;
}

template <typename AT, int NT>
int launch_fwd_mma(const msda_b200_desc* d, const WParams& w, const WinMaps& maps, cudaStream_t st) {
  const size_t smem = mma_smem_layout(w.win_bytes, w.guard_bytes, d->L, w.tqs).total;
  const long long blocks = (long long)w.k.B * w.k.num_tiles * (w.k.H / w.hpb);
  if (blocks > 0x7fffffffll) return msda_b200_internal_fail(MSDA_B200_ERR_UNSUPPORTED, "forward (window): grid too large");
#define MSDA_MMA_LAUNCH(LV)                                                                                           \
  {                                                                                                                   \
    auto kern = msda_fwd_mma_kernel<AT, LV, NT>;                                                                      \
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)            \
      return msda_b200_internal_fail(MSDA_B200_ERR_CUDA, "forward (window): cannot reserve shared memory");           \
    kern<<<(unsigned)blocks, NT, smem, st>>>(w, maps);                                                                \
  }
  switch (d->L) {
    case 1: MSDA_MMA_LAUNCH(1) break;
    case 2: MSDA_MMA_LAUNCH(2) break;
    case 3: MSDA_MMA_LAUNCH(3) break;
    case 4: MSDA_MMA_LAUNCH(4) break;
    default: return msda_b200_internal_fail(MSDA_B200_ERR_UNSUPPORTED, "forward (window): L > 4");
  }
#undef MSDA_MMA_LAUNCH
  return MSDA_B200_OK;
}

long long* g_win_dbg = nullptr;

}  // namespace

extern "C" {

// dev tool: device buffer of kWinDbgBlocks x kWinDbgSlots clock64() stamps (NULL switches the stamps off)
void msda_b200_internal_win_debug(long long* dev_buffer) { g_win_dbg = dev_buffer; }

// 1 when the window kernels cover this problem (the caller then launches through msda_b200_internal_win_forward)
int msda_b200_internal_win_applicable(const msda_b200_desc* d, const void* query_order) {
  // Opt-in (MSDA_B200_WINDOW=1): measured on B200 (profiles/r02_notes.md) the window-staged kernels are correct but,
  // at two 110 KB blocks per SM, slower than the per-corner gather of msda_b200.cu (0.40 / 0.61 ms vs 0.29 ms at
  // BASELINE config 2): the per-tile position / descriptor / TMA phases are not hidden behind the gather.
  if (d->flags & MSDA_B200_FLAG_NO_WINDOW) return 0;
  if (!env_int("MSDA_B200_WINDOW", 0)) return 0;
  if (!query_order || !d->tile_start || d->num_tiles <= 0) return 0;
  if (d->value_dtype != MSDA_B200_BF16 || d->D != 32 || d->P != kWinP || d->L > kWinL) return 0;
  if (d->max_tile > kWinTQ) return 0;
  for (int l = 0; l < d->L; ++l)
    if (d->spatial_shapes_hw[2 * l] > 8000 || d->spatial_shapes_hw[2 * l + 1] > 8000) return 0;
  return 1;
}

// kparams: the caller's KParams with tensor pointers and geometry filled in (fill_geometry); num_tiles is overwritten.
int msda_b200_internal_win_forward(const msda_b200_desc* d, const void* kparams, void* stream) {
  WParams w;
  memset(&w, 0, sizeof(w));
  memcpy(&w.k, kparams, sizeof(KParams));
  w.k.num_tiles = d->num_tiles;
  w.tile_start = d->tile_start;
  w.dbg = g_win_dbg;
  // descriptor row stride: = 1 mod 8, so the four weight packs a warp reads per level (16-byte slots, one per sample,
  // TQS * 16 bytes apart) fall into four different bank groups
  w.tqs = ((d->max_tile + 7) & ~7) + 1;
  // heads per block: as many as keep the grid at several waves of 2 blocks x 148 SMs
  w.hpb = 1;
  for (int c = 2; c <= d->H; c *= 2)
    if (d->H % c == 0 && (long long)d->B * d->num_tiles * (d->H / c) >= 6 * 296) w.hpb = c;
  if (env_int("MSDA_B200_WIN_HPB", 0) > 0 && d->H % env_int("MSDA_B200_WIN_HPB", 0) == 0) w.hpb = env_int("MSDA_B200_WIN_HPB", 0);
  const int kernel = env_int("MSDA_B200_WIN_KERNEL", 1);  // 1: tensor-core gather (default), 0: CUDA-core gather
  const int stage = env_int("MSDA_B200_WIN_STAGE", 0);    // CUDA-core kernel only: 0 TMA, 1 cp.async
  const int budget = 112 * 1024;  // two blocks per SM
  if (!plan_windows(d, w, budget, kernel == 1))
    return msda_b200_internal_fail(MSDA_B200_ERR_UNSUPPORTED, "forward (window): windows do not fit in shared memory");
  WinMaps maps;
  memset(&maps, 0, sizeof(maps));
  if ((kernel == 1 || stage == 0) && !encode_maps(d, w.k.value, w, maps))
    return msda_b200_internal_fail(MSDA_B200_ERR_CUDA, "forward (window): cuTensorMapEncodeTiled failed");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool abf = d->attn_dtype == MSDA_B200_BF16;
  int rc;
  if (kernel == 1) {
    if (env_int("MSDA_B200_WIN_NT", 256) == 512)
      rc = abf ? launch_fwd_mma<__nv_bfloat16, 512>(d, w, maps, st) : launch_fwd_mma<float, 512>(d, w, maps, st);
    else
      rc = abf ? launch_fwd_mma<__nv_bfloat16, 256>(d, w, maps, st) : launch_fwd_mma<float, 256>(d, w, maps, st);
  } else if (stage == 0) {
    rc = abf ? launch_fwd_win<__nv_bfloat16, 0>(d, w, maps, st) : launch_fwd_win<float, 0>(d, w, maps, st);
  } else {
    rc = abf ? launch_fwd_win<__nv_bfloat16, 1>(d, w, maps, st) : launch_fwd_win<float, 1>(d, w, maps, st);
  }
  return rc;
}

}  // extern "C"
