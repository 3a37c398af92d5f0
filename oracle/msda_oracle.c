/*
 * oracle/msda_oracle.c -- CPU restatement of multi-scale deformable attention.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path
 * (weed_instance_segmentation_b200/) may include, link or call this file; only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs use it, and only as the checker.
 *
 * What it restates (the algorithm lives in a third-party dependency of the
 * reference, not under /root/reference):
 *   transformers 5.5.0, models/mask2former/modeling_mask2former.py:798-837
 *   `multi_scale_deformable_attention(value, value_spatial_shapes,
 *   sampling_locations, attention_weights)`, reached from the reference at
 *   models/mask2former/train.py:196 (train), train.py:28 (val loss),
 *   models/metrics.py:56, models/mask2former/inference.py:27.
 *
 *   M2F:807      grid = 2*loc - 1
 *   M2F:822-824  grid_sample(bilinear, padding_mode="zeros", align_corners=False)
 *                => pixel coords  x = ((grid_x+1)*W - 1)/2 = loc_x*W - 0.5,
 *                                 y = ((grid_y+1)*H - 1)/2 = loc_y*H - 0.5,
 *                   four-corner bilinear, every out-of-range corner contributes 0
 *   M2F:806,815  level l occupies rows [start_l, start_l + H_l*W_l) of S,
 *                row-major (y*W_l + x)
 *   M2F:832-837  out[b,q,h*D+d] = sum_{l,p} attn[b,q,h,l,p] * sample(l,p)[d]
 *
 * Pinning: the reference repository has no tests or golden vectors (SURVEY.md
 * section 4).  This restatement is pinned against outputs of the reference's own
 * implementation (the function above, imported in the build container) stored
 * in tests/golden/ by tests/golden/make_golden.py, and re-checked live against
 * that function wherever `transformers` is importable (tests/test_oracle.py).
 *
 * Layouts (all contiguous, C order):
 *   value  (B, S, H, D)      shapes_hw (L, 2) int32 = (H_l, W_l)
 *   loc    (B, Q, H, L, P, 2) last dim (x, y), normalised
 *   attn   (B, Q, H, L, P)    level_start (L,) int64
 *   out    (B, Q, H*D)
 *
 * The backward is the analytic gradient of the forward (what autograd produces
 * through M2F:798-837):
 *   grad_value[corner] += attn * w_corner * grad_out
 *   grad_attn          = <grad_out, bilinear sample>
 *   grad_loc.x         = W_l * attn * <grad_out, d sample / d px>   (y likewise with H_l)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MSDA_ORACLE_IMPL(REAL, SUFFIX)                                                      \
  int msda_oracle_forward_##SUFFIX(const REAL* value, const int32_t* shapes_hw,             \
                                   const int64_t* level_start, const REAL* loc,             \
                                   const REAL* attn, REAL* out, int B, int S, int Q, int H, \
                                   int D, int L, int P) {                                   \
    if (B < 0 || S < 0 || Q < 0 || H <= 0 || D <= 0 || L <= 0 || P <= 0) return 1;          \
    for (int l = 0; l < L; ++l) {                                                           \
      int64_t n = (int64_t)shapes_hw[2 * l] * shapes_hw[2 * l + 1];                         \
      if (shapes_hw[2 * l] <= 0 || shapes_hw[2 * l + 1] <= 0) return 2;                     \
      if (level_start[l] < 0 || level_start[l] + n > S) return 3;                           \
    }                                                                                       \
    for (int b = 0; b < B; ++b)                                                             \
      for (int q = 0; q < Q; ++q)                                                           \
        for (int h = 0; h < H; ++h) {                                                       \
          REAL* o = out + (((int64_t)b * Q + q) * H + h) * D;                               \
          for (int d = 0; d < D; ++d) o[d] = (REAL)0;                                       \
          for (int l = 0; l < L; ++l) {                                                     \
            const int Hl = shapes_hw[2 * l], Wl = shapes_hw[2 * l + 1];                     \
            const REAL* vl = value + ((int64_t)b * S + level_start[l]) * H * D;             \
            for (int p = 0; p < P; ++p) {                                                   \
              const int64_t si = ((((int64_t)b * Q + q) * H + h) * L + l) * P + p;          \
              const REAL a = attn[si];                                                      \
              /* M2F:807 then ATen grid_sampler_unnormalize(align_corners=False) */        \
              const REAL gx = (REAL)2 * loc[2 * si] - (REAL)1;                              \
              const REAL gy = (REAL)2 * loc[2 * si + 1] - (REAL)1;                          \
              const REAL px = ((gx + (REAL)1) * (REAL)Wl - (REAL)1) / (REAL)2;              \
              const REAL py = ((gy + (REAL)1) * (REAL)Hl - (REAL)1) / (REAL)2;              \
              /* Every corner out of range (also NaN). px == -1 is NOT skipped: its     */ \
              /* right-hand corner is pixel 0 with weight 0 but a non-zero d/dpx (ATen    */ \
              /* does the same: floor(-1.0) = -1, corner x0+1 = 0 is in bounds).          */ \
              if (!(px > (REAL)-2 && px < (REAL)Wl + 1 && py > (REAL)-2 && py < (REAL)Hl + 1)) \
                continue;                                                                   \
              const REAL fx = (REAL)floor((double)px), fy = (REAL)floor((double)py);        \
              const int x0 = (int)fx, y0 = (int)fy;                                         \
              const REAL lx = px - fx, ly = py - fy;                                        \
              const REAL wgt[4] = {((REAL)1 - lx) * ((REAL)1 - ly), lx * ((REAL)1 - ly),    \
                                   ((REAL)1 - lx) * ly, lx * ly};                           \
              for (int c = 0; c < 4; ++c) {                                                 \
                const int xx = x0 + (c & 1), yy = y0 + (c >> 1);                            \
                if (xx < 0 || xx >= Wl || yy < 0 || yy >= Hl) continue;                     \
                const REAL* v = vl + (((int64_t)yy * Wl + xx) * H + h) * D;                 \
                const REAL w = a * wgt[c];                                                  \
                for (int d = 0; d < D; ++d) o[d] += w * v[d];                               \
              }                                                                             \
            }                                                                               \
          }                                                                                 \
        }                                                                                   \
    return 0;                                                                               \
  }                                                                                         \
                                                                                            \
  int msda_oracle_backward_##SUFFIX(const REAL* value, const int32_t* shapes_hw,            \
                                    const int64_t* level_start, const REAL* loc,            \
                                    const REAL* attn, const REAL* grad_out,                 \
                                    REAL* grad_value, REAL* grad_loc, REAL* grad_attn,      \
                                    int B, int S, int Q, int H, int D, int L, int P) {      \
    if (B < 0 || S < 0 || Q < 0 || H <= 0 || D <= 0 || L <= 0 || P <= 0) return 1;          \
    for (int l = 0; l < L; ++l) {                                                           \
      int64_t n = (int64_t)shapes_hw[2 * l] * shapes_hw[2 * l + 1];                         \
      if (shapes_hw[2 * l] <= 0 || shapes_hw[2 * l + 1] <= 0) return 2;                     \
      if (level_start[l] < 0 || level_start[l] + n > S) return 3;                           \
    }                                                                                       \
    memset(grad_value, 0, sizeof(REAL) * (size_t)B * S * H * D);                            \
    for (int b = 0; b < B; ++b)                                                             \
      for (int q = 0; q < Q; ++q)                                                           \
        for (int h = 0; h < H; ++h) {                                                       \
          const REAL* go = grad_out + (((int64_t)b * Q + q) * H + h) * D;                   \
          for (int l = 0; l < L; ++l) {                                                     \
            const int Hl = shapes_hw[2 * l], Wl = shapes_hw[2 * l + 1];                     \
            const int64_t lvl = ((int64_t)b * S + level_start[l]) * H * D;                  \
            for (int p = 0; p < P; ++p) {                                                   \
              const int64_t si = ((((int64_t)b * Q + q) * H + h) * L + l) * P + p;          \
              const REAL a = attn[si];                                                      \
              grad_attn[si] = (REAL)0;                                                      \
              grad_loc[2 * si] = (REAL)0;                                                   \
              grad_loc[2 * si + 1] = (REAL)0;                                               \
              const REAL gx = (REAL)2 * loc[2 * si] - (REAL)1;                              \
              const REAL gy = (REAL)2 * loc[2 * si + 1] - (REAL)1;                          \
              const REAL px = ((gx + (REAL)1) * (REAL)Wl - (REAL)1) / (REAL)2;              \
              const REAL py = ((gy + (REAL)1) * (REAL)Hl - (REAL)1) / (REAL)2;              \
              if (!(px > (REAL)-2 && px < (REAL)Wl + 1 && py > (REAL)-2 && py < (REAL)Hl + 1)) \
                continue;                                                                   \
              const REAL fx = (REAL)floor((double)px), fy = (REAL)floor((double)py);        \
              const int x0 = (int)fx, y0 = (int)fy;                                         \
              const REAL lx = px - fx, ly = py - fy;                                        \
              const REAL wgt[4] = {((REAL)1 - lx) * ((REAL)1 - ly), lx * ((REAL)1 - ly),    \
                                   ((REAL)1 - lx) * ly, lx * ly};                           \
              /* d wgt / d px and d wgt / d py for corners (0,0) (1,0) (0,1) (1,1) */       \
              const REAL dwx[4] = {-((REAL)1 - ly), ((REAL)1 - ly), -ly, ly};               \
              const REAL dwy[4] = {-((REAL)1 - lx), -lx, ((REAL)1 - lx), lx};               \
              REAL ga = (REAL)0, gpx = (REAL)0, gpy = (REAL)0;                              \
              for (int c = 0; c < 4; ++c) {                                                 \
                const int xx = x0 + (c & 1), yy = y0 + (c >> 1);                            \
                if (xx < 0 || xx >= Wl || yy < 0 || yy >= Hl) continue;                     \
                const int64_t vo = lvl + (((int64_t)yy * Wl + xx) * H + h) * D;             \
                REAL dot = (REAL)0;                                                         \
                const REAL w = a * wgt[c];                                                  \
                for (int d = 0; d < D; ++d) {                                               \
                  dot += go[d] * value[vo + d];                                             \
                  grad_value[vo + d] += w * go[d];                                          \
                }                                                                           \
                ga += wgt[c] * dot;                                                         \
                gpx += dwx[c] * dot;                                                        \
                gpy += dwy[c] * dot;                                                        \
              }                                                                             \
              grad_attn[si] = ga;                                                           \
              grad_loc[2 * si] = (REAL)Wl * a * gpx;                                        \
              grad_loc[2 * si + 1] = (REAL)Hl * a * gpy;                                    \
            }                                                                               \
          }                                                                                 \
        }                                                                                   \
    return 0;                                                                               \
  }

MSDA_ORACLE_IMPL(float, f32)
MSDA_ORACLE_IMPL(double, f64)
