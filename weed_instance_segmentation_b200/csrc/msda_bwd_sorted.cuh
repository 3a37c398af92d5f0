// msda_bwd_sorted.cuh -- backward v2: pixel-sorted accumulation ("level-sorted accumulation" of the north star).
// Included by msda_b200.cu inside its anonymous namespace (uses KParams, Level, Axis, Vec16, red_add_*).
//
// Why: v1 issues one fp32 reduction per bilinear corner (66 M corners x 128 B at BASELINE config 2) and is bound
// by L2 reduction throughput (measured 6.2 TB/s); shared-memory fp32 atomics are a CAS loop on sm_100a, so the
// aggregation on chip is done by *ownership* instead:
//
//   per thread block = (batch, head, tile of TQ queries), per level:
//     a  sample descriptors (slot weights, derivative codes, clamped base pixel) -> shared memory; bounding box
//        of the touched pixels with warp reductions + native integer shared atomics
//     b  window = bounding box (or a CAP-pixel rectangle around the mean if the box is larger)
//     c  histogram of the TQ*P*4 corner contributions over window pixels (ATOMS.ADD, integer)
//     d  block-wide exclusive scan -> segment starts
//     e  fill: contributions sorted by target pixel (counting sort), 8-byte entries {pixel | id, weight}
//     f  pull: lane groups walk equal shares of the sorted list; per entry one LDS.128 of the staged grad_out
//        row, FMA into a register accumulator (grad_value) and a dot product with the pixel's value row
//        (for grad_loc / grad_attn); ONE reduction to global memory per run of equal pixels, and the value row
//        is read once per run instead of once per corner
//     g  contributions outside the window take the v1 route (direct reduction), so any input is handled
//     h  one thread per sample folds its four dots into grad_attn / grad_loc
//
// Results are identical to v1 up to fp32 summation order.

template <int P>
struct Log2P;
template <> struct Log2P<1> { static constexpr int v = 0; };
template <> struct Log2P<2> { static constexpr int v = 1; };
template <> struct Log2P<4> { static constexpr int v = 2; };
template <> struct Log2P<8> { static constexpr int v = 3; };

struct SortedSmem {
  // byte offsets into dynamic shared memory
  size_t w, ent, dot, a, xy, go, cnt, fb, misc, stat, total;
};

template <int NT, int TQ, int P, int CAP, int LPP>
__host__ __device__ inline SortedSmem sorted_smem_layout() {
  constexpr int NS = TQ * P, NC = NS * 4;
  SortedSmem s;
  size_t o = 0;
  s.w = o;    o += sizeof(float4) * NS;
  s.go = o;   o += sizeof(uint4) * (TQ + 1) * LPP;  // + one all-zero row, read by the padding entries of the pull
  s.ent = o;  o += sizeof(int2) * NC;
  s.dot = o;  o += sizeof(float) * NC;
  s.a = o;    o += sizeof(float) * NS;
  s.xy = o;   o += sizeof(int) * NS;  // x | y << 12 | derivative codes << 24 (level extents <= 4095)
  s.cnt = o;  o += sizeof(int) * (CAP + 4);
  s.fb = s.ent + sizeof(int2) * NC;  // fallback ids grow downwards from the end of the entry array (E + nfb <= NC)
  o = (o + 15) & ~size_t(15);
  s.misc = o; o += sizeof(int) * 64;
  s.stat = o; o += sizeof(float) * 3 * TQ;  // fused prologue: softmax max, 1/sum, sum_k a_k*ga_k per query
  s.total = o;
  return s;
}

// misc slots
enum { MI_MINX = 0, MI_MAXX, MI_MINY, MI_MAXY, MI_SUMX, MI_SUMY, MI_NACT, MI_FBN, MI_TOTAL, MI_WSUM = 16 };

__device__ __forceinline__ float gdec(int code, int k) { return (float)(((code >> (2 * k)) & 3) - 1); }

template <typename VT, typename AT, int D, int NT, int TQ, int P, int CAP, int ACC, bool FUSED, int PULL_LANES>
__global__ void __launch_bounds__(NT, 1024 / NT) msda_bwd_sorted_kernel(const __grid_constant__ KParams p) {
  constexpr int VEC = Vec16<VT>::N;
  constexpr int LPP = D / VEC;
  constexpr int G = NT / LPP;
  constexpr int NS = TQ * P, NC = NS * 4;
  constexpr int LP2 = Log2P<P>::v;
  constexpr int WMAX = 256;  // window edge limit (8 bits per coordinate in the sort key)
  static_assert(CAP % NT == 0 && (CAP / NT) % 4 == 0, "scan assumes CAP/NT is a multiple of 4");
  static_assert(NC <= 65536, "entry ids are 16 bit");
  static_assert(NS % NT == 0, "every thread owns NS/NT samples per level");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const SortedSmem lay = sorted_smem_layout<NT, TQ, P, CAP, LPP>();
  float4* s_w = reinterpret_cast<float4*>(smem_raw + lay.w);
  uint4* s_go = reinterpret_cast<uint4*>(smem_raw + lay.go);
  int2* s_ent = reinterpret_cast<int2*>(smem_raw + lay.ent);
  float* s_dot = reinterpret_cast<float*>(smem_raw + lay.dot);
  float* s_a = reinterpret_cast<float*>(smem_raw + lay.a);
  int* s_xy = reinterpret_cast<int*>(smem_raw + lay.xy);
  int* s_cnt = reinterpret_cast<int*>(smem_raw + lay.cnt);
  unsigned short* s_fb_end = reinterpret_cast<unsigned short*>(smem_raw + lay.fb);
  int* s_misc = reinterpret_cast<int*>(smem_raw + lay.misc);
  float* s_max = reinterpret_cast<float*>(smem_raw + lay.stat);
  float* s_inv = s_max + TQ;
  float* s_dsum = s_inv + TQ;

  int b, tile, h;
  decode_block(p, b, tile, h);
  const int q0 = tile * TQ;
  const int nq = min(TQ, p.Q - q0);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = tid / LPP, c = tid % LPP;
  const unsigned gmask = (LPP >= 32) ? 0xffffffffu : (((1u << LPP) - 1u) << ((lane / LPP) * LPP));
  const uint4* vb = reinterpret_cast<const uint4*>(p.value) + (long long)b * p.batch_stride16 + c;
  const long long acc_base = (long long)b * p.batch_stride16;

  // grad_out rows of the tile, staged once (zero rows for the ragged tail)
  for (int i = tid; i < TQ * LPP; i += NT) {
    const int ql = i / LPP, cc = i % LPP;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (ql < nq) {
      const int q = p.q_order ? p.q_order[q0 + ql] : q0 + ql;
      v = ldg16(reinterpret_cast<const uint4*>(p.grad_out) + (((long long)b * p.Q + q) * p.H + h) * LPP + cc);
    }
    s_go[i] = v;
  }
  if (tid < LPP) s_go[TQ * LPP + tid] = make_uint4(0u, 0u, 0u, 0u);

  // per-thread samples: si = r*NT + tid; their loc / attn are prefetched one level ahead
  constexpr int SPT = (NS + NT - 1) / NT;
  float2 pre_loc[SPT];
  float pre_a[SPT];
  auto fetch_level = [&](int l) {
#pragma unroll
    for (int r = 0; r < SPT; ++r) {
      const int si = r * NT + tid;
      const int ql = si >> LP2, pt = si & (P - 1);
      pre_loc[r] = make_float2(0.f, 0.f);
      pre_a[r] = 0.f;
      if (l < p.L && si < NS && ql < nq) {
        const int q = p.q_order ? p.q_order[q0 + ql] : q0 + ql;
        const long long gi = (((long long)b * p.Q + q) * p.H + h) * p.LP + l * P + pt;
        if (FUSED) {
          pre_loc[r] = fused_loc<AT>(p, gi, ((long long)b * p.Q + q) * p.L + l, q, p.lv[l]);
          pre_a[r] = to_float<AT>(reinterpret_cast<const AT*>(p.logits)[gi]);  // raw logit; softmax applied at use
        } else {
          pre_loc[r] = __ldg(reinterpret_cast<const float2*>(p.loc) + gi);
          pre_a[r] = to_float<AT>(reinterpret_cast<const AT*>(p.attn)[gi]);
        }
      }
    }
  };
  fetch_level(0);
  if (FUSED) {
    // softmax statistics of every (query, head) row of the tile (four lanes per row); visible after the first
    // barrier of the level loop
    for (int qb = 0; qb < TQ; qb += NT / 4) {
      const int ql = qb + (tid >> 2);
      const bool valid = ql < nq;
      const int q = valid ? (p.q_order ? p.q_order[q0 + ql] : q0 + ql) : 0;
      float mx, inv;
      softmax_stats_x4<AT>(reinterpret_cast<const AT*>(p.logits) + (((long long)b * p.Q + q) * p.H + h) * p.LP, p.LP,
                           tid & 3, valid, mx, inv);
      if ((tid & 3) == 0) { s_max[ql] = valid ? mx : 0.f; s_inv[ql] = inv; s_dsum[ql] = 0.f; }
    }
  }

  for (int l = 0; l < p.L; ++l) {
    const Level lv = p.lv[l];
    const int dxs = lv.W > 1 ? 1 : 0, dys = lv.H > 1 ? 1 : 0;
    if (tid < 16) {
      int init = 0;
      if (tid == MI_MINX || tid == MI_MINY) init = 0x7fffffff;
      if (tid == MI_MAXX || tid == MI_MAXY) init = -1;
      s_misc[tid] = init;
    }
    for (int i = tid; i < CAP + 4; i += NT) s_cnt[i] = 0;
    __syncthreads();

    // ---- a: descriptors + bounding box (loc / attn of this level were fetched during the previous level)
#pragma unroll
    for (int r = 0; r < SPT; ++r) {
      const int si = r * NT + tid;
      const int ql = si >> LP2;
      float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
      int gcode = 0x55;  // code 1 == derivative 0
      float a = 0.f;
      bool active = false;
      int xb = 0, yb = 0;
      if (si < NS && ql < nq) {
        a = FUSED ? expf(pre_a[r] - s_max[ql]) * s_inv[ql] : pre_a[r];
        const Axis ax = axis_setup(pre_loc[r].x, lv.W), ay = axis_setup(pre_loc[r].y, lv.H);
        if (ax.ok && ay.ok) {
          w = make_float4(ax.s0, ax.s1, ay.s0, ay.s1);
          gcode = ((int)ax.g0 + 1) | (((int)ax.g1 + 1) << 2) | (((int)ay.g0 + 1) << 4) | (((int)ay.g1 + 1) << 6);
          const bool xa = ax.s0 != 0.f || ax.s1 != 0.f || ax.g0 != 0.f || ax.g1 != 0.f;
          const bool ya = ay.s0 != 0.f || ay.s1 != 0.f || ay.g0 != 0.f || ay.g1 != 0.f;
          active = xa && ya;
        }
        xb = ax.base; yb = ay.base;
      }
      if (si < NS) {
        s_w[si] = w; s_a[si] = a; s_xy[si] = xb | (yb << 12) | (gcode << 24);
      }
      const unsigned act = __ballot_sync(0xffffffffu, active);
      if (act) {
        const int mnx = __reduce_min_sync(0xffffffffu, active ? xb : 0x7fffffff);
        const int mxx = __reduce_max_sync(0xffffffffu, active ? xb + dxs : -1);
        const int mny = __reduce_min_sync(0xffffffffu, active ? yb : 0x7fffffff);
        const int mxy = __reduce_max_sync(0xffffffffu, active ? yb + dys : -1);
        const int sx = __reduce_add_sync(0xffffffffu, active ? xb : 0);
        const int sy = __reduce_add_sync(0xffffffffu, active ? yb : 0);
        if (lane == 0) {
          atomicMin(&s_misc[MI_MINX], mnx); atomicMax(&s_misc[MI_MAXX], mxx);
          atomicMin(&s_misc[MI_MINY], mny); atomicMax(&s_misc[MI_MAXY], mxy);
          atomicAdd(&s_misc[MI_SUMX], sx); atomicAdd(&s_misc[MI_SUMY], sy);
          atomicAdd(&s_misc[MI_NACT], __popc(act));
        }
      }
    }
    fetch_level(l + 1);  // in flight during the histogram / sort / pull of this level
    __syncthreads();

    // ---- b: window (every thread computes the same rectangle)
    int x0 = 0, y0 = 0, ww = 0, wh = 0;
    {
      const int nact = s_misc[MI_NACT];
      if (nact > 0) {
        const int mnx = s_misc[MI_MINX], mxx = s_misc[MI_MAXX], mny = s_misc[MI_MINY], mxy = s_misc[MI_MAXY];
        const int bw = mxx - mnx + 1, bh = mxy - mny + 1;
        if (bw <= WMAX && bh <= WMAX && bw * bh <= CAP) {
          x0 = mnx; y0 = mny; ww = bw; wh = bh;
        } else {
          ww = min(bw, 64);
          wh = min(min(bh, CAP / ww), WMAX);
          const int cx = s_misc[MI_SUMX] / nact, cy = s_misc[MI_SUMY] / nact;
          x0 = min(max(cx - ww / 2, mnx), mxx - ww + 1);
          y0 = min(max(cy - wh / 2, mny), mxy - wh + 1);
        }
      }
    }

    // ---- c: histogram over window pixels; contributions outside the window go to the fallback list.
    // One thread per *sample*: its four corners share the descriptor loads; the window code of every corner
    // ((py << 8) | px, -1 = inactive, -2 = fallback) stays in registers for the fill pass (e).
    int code[SPT][4];
#pragma unroll
    for (int r = 0; r < SPT; ++r) {
      const int si = r * NT + tid;
      const float4 w = s_w[si];
      const int xy = s_xy[si];
      const int gcode = xy >> 24;
      const bool xa0 = w.x != 0.f || (gcode & 3) != 1, xa1 = w.y != 0.f || ((gcode >> 2) & 3) != 1;
      const bool ya0 = w.z != 0.f || ((gcode >> 4) & 3) != 1, ya1 = w.w != 0.f || ((gcode >> 6) & 3) != 1;
      const int bx = (xy & 0xfff) - x0, by = ((xy >> 12) & 0xfff) - y0;
      *reinterpret_cast<float4*>(&s_dot[si * 4]) = make_float4(0.f, 0.f, 0.f, 0.f);  // active corners are overwritten in f / g
#pragma unroll
      for (int cn = 0; cn < 4; ++cn) {
        const bool active = ((cn & 1) ? xa1 : xa0) && ((cn & 2) ? ya1 : ya0);
        const int px = bx + ((cn & 1) ? dxs : 0), py = by + ((cn & 2) ? dys : 0);
        int cd = -1;
        if (active) {
          if ((unsigned)px < (unsigned)ww && (unsigned)py < (unsigned)wh) {
            cd = (py << 8) | px;
            atomicAdd(&s_cnt[py * ww + px], 1);
          } else {
            cd = -2;
            s_fb_end[-1 - atomicAdd(&s_misc[MI_FBN], 1)] = (unsigned short)(si * 4 + cn);
          }
        }
        code[r][cn] = cd;
      }
    }
    __syncthreads();

    // ---- d: exclusive scan of s_cnt[0..CAP)
    {
      constexpr int PER = CAP / NT;
      int v[PER];
      int sum = 0;
#pragma unroll
      for (int k = 0; k < PER; k += 4) {
        const int4 t = *reinterpret_cast<const int4*>(&s_cnt[tid * PER + k]);
        v[k] = t.x; v[k + 1] = t.y; v[k + 2] = t.z; v[k + 3] = t.w;
        sum += t.x + t.y + t.z + t.w;
      }
      int incl = sum;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += t;
      }
      if (lane == 31) s_misc[MI_WSUM + warp] = incl;
      __syncthreads();
      if (warp == 0) {
        const int t = lane < NT / 32 ? s_misc[MI_WSUM + lane] : 0;
        int inc2 = t;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
          const int u = __shfl_up_sync(0xffffffffu, inc2, off);
          if (lane >= off) inc2 += u;
        }
        if (lane < NT / 32) s_misc[MI_WSUM + lane] = inc2 - t;
        if (lane == NT / 32 - 1) s_misc[MI_TOTAL] = inc2;
      }
      __syncthreads();
      int base = s_misc[MI_WSUM + warp] + incl - sum;
#pragma unroll
      for (int k = 0; k < PER; ++k) {
        const int t = v[k];
        v[k] = base;
        base += t;
      }
#pragma unroll
      for (int k = 0; k < PER; k += 4)
        *reinterpret_cast<int4*>(&s_cnt[tid * PER + k]) = make_int4(v[k], v[k + 1], v[k + 2], v[k + 3]);
    }
    __syncthreads();

    // ---- e: fill (counting sort by window pixel), from the codes kept in registers
#pragma unroll
    for (int r = 0; r < SPT; ++r) {
      const int si = r * NT + tid;
      const float4 w = s_w[si];
      const float a = s_a[si];
      const float wy0 = a * w.z, wy1 = a * w.w;
#pragma unroll
      for (int cn = 0; cn < 4; ++cn) {
        const int cd = code[r][cn];
        if (cd >= 0) {
          const int slot = atomicAdd(&s_cnt[(cd >> 8) * ww + (cd & 0xff)], 1);
          s_ent[slot] = make_int2((cd << 16) | (si * 4 + cn),
                                  __float_as_int(((cn & 2) ? wy1 : wy0) * ((cn & 1) ? w.y : w.x)));
        }
      }
    }
    __syncthreads();

    // ---- f: pull over the sorted list: PL lanes share one entry (each lane owns CH 16-byte chunks of the
    // row); every lane group walks `per` consecutive entries -- the same trip count for all groups, so the
    // warp stays convergent and the dot reduction can use full-mask shuffles
    {
      constexpr int PL = (LPP >= PULL_LANES) ? PULL_LANES : LPP;  // lanes per entry
      constexpr int CH = LPP / PL;                               // 16-byte chunks per lane
      constexpr int GP = NT / PL;                                // entries in flight per block
      const int pg = tid / PL, pc = (tid % PL) * CH;
      const int E = s_misc[MI_TOTAL];
      // share length per lane group, forced odd: the groups of a warp then read their 8-byte entries from
      // distinct banks (an even share -- 2048 entries / 64 groups = 32 is the common case -- puts all of them
      // 256 B apart, an 8-way bank conflict on every entry load)
      const int per = ((E + GP - 1) / GP) | 1;
      const int e0 = pg * per;
      const int e_end = min(e0 + per, E);
      // window pixel (px, py) -> 16-byte-unit offset inside the batch item: base0 + py * sy + px * sx
      const int sx = p.H * LPP, sy = lv.W * sx;
      const int base0 = ((lv.start + y0 * lv.W + x0) * p.H + h) * LPP + pc;
      constexpr int kChunkBytes = VEC * (ACC == 0 ? (int)sizeof(float) : (int)sizeof(VT));
      char* const acc_ptr = reinterpret_cast<char*>(p.grad_value_acc) + acc_base * (long long)kChunkBytes;
      const uint4* const vrow = reinterpret_cast<const uint4*>(p.value) + acc_base;
      const uint4* const go_lane = s_go + pc;
      int cur = -1, cur_off = 0;
      float acc[CH][VEC], vf[CH][VEC];
#pragma unroll
      for (int u = 0; u < CH; ++u)
#pragma unroll
        for (int j = 0; j < VEC; ++j) { acc[u][j] = 0.f; vf[u][j] = 0.f; }
      auto flush = [&]() {
#pragma unroll
        for (int u = 0; u < CH; ++u) {
          char* dst = acc_ptr + (long long)(cur_off + u) * kChunkBytes;
          if (ACC == 0) {
#pragma unroll
            for (int j = 0; j < VEC; j += 4)
              red_add_f32x4(reinterpret_cast<float*>(dst) + j, acc[u][j], acc[u][j + 1], acc[u][j + 2], acc[u][j + 3]);
          } else {
            const uint4 pk = Vec16<VT>::pack(acc[u]);
            red_add_bf16x8(dst, pk.x, pk.y, pk.z, pk.w);
          }
        }
      };
      for (int k = 0; k < per; ++k) {
        const int e = e0 + k;
        const bool valid = e < e_end;
        // padding: same pixel, weight 0, the all-zero grad_out row (0 * a non-finite grad_out would poison the run), dot discarded
        int2 en = make_int2((cur << 16) | (TQ << (2 + LP2)), 0);
        if (valid) en = s_ent[e];
        const int pix = (int)((unsigned)en.x >> 16);
        const int id = en.x & 0xffff;
        const float wgt = __int_as_float(en.y);
        if (valid && pix != cur) {
          if (cur >= 0) flush();
          cur = pix;
          cur_off = base0 + (pix >> 8) * sy + (pix & 0xff) * sx;
#pragma unroll
          for (int u = 0; u < CH; ++u) {
            Vec16<VT>::unpack(ldg16(vrow + cur_off + u), vf[u]);
#pragma unroll
            for (int j = 0; j < VEC; ++j) acc[u][j] = 0.f;
          }
        }
        float d0 = 0.f, d1 = 0.f;  // even / odd channels (packed FFMA2)
#pragma unroll
        for (int u = 0; u < CH; ++u) {
          float gf[VEC];
          Vec16<VT>::unpack(go_lane[(id >> (2 + LP2)) * LPP + u], gf);
#pragma unroll
          for (int j = 0; j < VEC; j += 2) {
            fma2_scalar(acc[u][j], acc[u][j + 1], wgt, gf[j], gf[j + 1]);
            fma2_pair(d0, d1, gf[j], gf[j + 1], vf[u][j], vf[u][j + 1]);
          }
        }
        float d = d0 + d1;
#pragma unroll
        for (int o = 1; o < PL; o <<= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
        if (valid && pc == 0) s_dot[id] = d;
      }
      if (cur >= 0) flush();
    }

    // ---- g: contributions outside the window: direct reduction (v1 route)
    {
      const int nfb = s_misc[MI_FBN];
      for (int k = g; k < nfb; k += G) {
        const int id = s_fb_end[-1 - k];
        const int si = id >> 2, cn = id & 3;
        const float4 w = s_w[si];
        const float wgt = s_a[si] * ((cn & 2) ? w.w : w.z) * ((cn & 1) ? w.y : w.x);
        const int xy = s_xy[si];
        const int x = (xy & 0xfff) + ((cn & 1) ? dxs : 0), y = ((xy >> 12) & 0xfff) + ((cn & 2) ? dys : 0);
        const long long off = (long long)((lv.start + y * lv.W + x) * p.H + h) * LPP;
        float vf[VEC], gf[VEC];
        Vec16<VT>::unpack(ldg16(vb + off), vf);
        Vec16<VT>::unpack(s_go[(id >> (2 + LP2)) * LPP + c], gf);
        float d = 0.f;
#pragma unroll
        for (int j = 0; j < VEC; ++j) d = fmaf(gf[j], vf[j], d);
#pragma unroll
        for (int o = 1; o < LPP; o <<= 1) d += __shfl_xor_sync(gmask, d, o);
        if (c == 0) s_dot[id] = d;
        if (wgt != 0.f) {
          if (ACC == 0) {
            float* dst = reinterpret_cast<float*>(p.grad_value_acc) + (acc_base + off + c) * VEC;
#pragma unroll
            for (int j = 0; j < VEC; j += 4)
              red_add_f32x4(dst + j, wgt * gf[j], wgt * gf[j + 1], wgt * gf[j + 2], wgt * gf[j + 3]);
          } else {
            float t[VEC];
#pragma unroll
            for (int j = 0; j < VEC; ++j) t[j] = wgt * gf[j];
            const uint4 pk = Vec16<VT>::pack(t);
            red_add_bf16x8(reinterpret_cast<VT*>(p.grad_value_acc) + (acc_base + off + c) * VEC, pk.x, pk.y, pk.z, pk.w);
          }
        }
      }
    }
    __syncthreads();

    // ---- h: per-sample gradients
    for (int si = tid; si < NS; si += NT) {
      const int ql = si >> LP2, pt = si & (P - 1);
      const bool valid = ql < nq;
      const int q = valid ? (p.q_order ? p.q_order[q0 + ql] : q0 + ql) : 0;
      const long long gi = (((long long)b * p.Q + q) * p.H + h) * p.LP + l * P + pt;
      const float4 d = *reinterpret_cast<const float4*>(&s_dot[si * 4]);
      const float4 w = s_w[si];
      const int gcode = s_xy[si] >> 24;
      const float a = s_a[si];
      const float gl = gdec(gcode, 0), gr = gdec(gcode, 1), gt = gdec(gcode, 2), gb = gdec(gcode, 3);
      const float top_s = w.x * d.x + w.y * d.y, bot_s = w.x * d.z + w.y * d.w;
      const float top_g = gl * d.x + gr * d.y, bot_g = gl * d.z + gr * d.w;
      const float g_attn = w.z * top_s + w.w * bot_s;
      const float g_px = w.z * top_g + w.w * bot_g, g_py = gt * top_s + gb * bot_s;
      if (!FUSED) {
        if (valid) {
          reinterpret_cast<AT*>(p.grad_attn)[gi] = from_float<AT>(g_attn);
          reinterpret_cast<float2*>(p.grad_loc)[gi] = make_float2((float)lv.W * a * g_px, (float)lv.H * a * g_py);
        }
      } else {
        if (valid) {
          store_pair<AT>(p.grad_offsets, gi, a * g_px, a * g_py);
          reinterpret_cast<AT*>(p.grad_logits)[gi] = from_float<AT>(g_attn);  // parked; finished after the last level
        }
        float t = valid ? a * g_attn : 0.f;  // the P samples of one (query, level) sit in P adjacent lanes
#pragma unroll
        for (int o = 1; o < P; o <<= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (valid && pt == 0) s_dsum[ql] += t;  // one writer per query and level; levels are separated by barriers
      }
    }
    __syncthreads();
  }

  if (FUSED) {
    // softmax backward over the L*P logits of each (query, head): g_j = a_j * (ga_j - sum_k a_k ga_k)
    for (int i = tid; i < nq * p.LP; i += NT) {
      const int ql = i / p.LP, sidx = i - ql * p.LP;
      const int q = p.q_order ? p.q_order[q0 + ql] : q0 + ql;
      const long long gi = (((long long)b * p.Q + q) * p.H + h) * p.LP + sidx;
      const float a = expf(to_float<AT>(reinterpret_cast<const AT*>(p.logits)[gi]) - s_max[ql]) * s_inv[ql];
      const float ga = to_float<AT>(reinterpret_cast<const AT*>(p.grad_logits)[gi]);
      reinterpret_cast<AT*>(p.grad_logits)[gi] = from_float<AT>(a * (ga - s_dsum[ql]));
    }
  }
}
