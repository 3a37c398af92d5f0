// Batched bilinear point sampling for the loss / matcher path.
//
// The reference samples masks at random points through `sample_point` (M2F:245-274: grid_sample, bilinear,
// zeros padding, align_corners=False) once per image in the Hungarian matcher (M2F:455-474) and three times per
// decoder layer in the mask loss (M2F:619-631, :718-737), each call on freshly gathered / padded copies of the masks.
// Here one launch samples any number of ROWS: a row is one 2-D plane (a pointer, so planes stay where they are --
// no gather, no padding, no dtype upcast), its extent, and the row of the coordinate table it reads (several planes
// may share one point set, as in the matcher).  The backward scatters into fp32 planes with atomics.
//
// Same bilinear primitive as the MSDeformAttn kernels: pix = ((2c - 1 + 1) * n - 1) / 2 evaluated in ATen's order
// without FMA contraction, corners outside the plane contribute zero, NaN coordinates contribute zero.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>

#include "msda_b200.h"

extern "C" int msda_b200_internal_fail(int code, const char* msg);  // msda_b200.cu: sets msda_b200_last_error()
extern "C" void msda_b200_internal_count_launch(void);               // msda_b200.cu: launch counter

namespace {

constexpr int kPsThreads = 256;

struct Corner {
  int i0;        // floor(pix)
  float w0, w1;  // weights of i0 and i0 + 1 (zero when outside [0, n))
};

__device__ __forceinline__ Corner corner_setup(float c, int n) {
  const float g = __fadd_rn(__fmul_rn(2.f, c), -1.f);
  const float pix = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(g, 1.f), (float)n), -1.f), 0.5f);
  Corner r;
  const bool ok = (pix > -2.f) && (pix < (float)(n + 1));  // false for NaN
  const float fl = floorf(pix);
  r.i0 = ok ? __float2int_rd(pix) : -4;
  const float l = pix - fl;
  r.w0 = (r.i0 >= 0 && r.i0 < n) ? 1.f - l : 0.f;
  r.w1 = (r.i0 + 1 >= 0 && r.i0 + 1 < n) ? l : 0.f;
  return r;
}

__device__ __forceinline__ float load_plane(const void* plane, int dtype, long long idx) {
  if (dtype == MSDA_B200_U8) return (float)__ldg(reinterpret_cast<const unsigned char*>(plane) + idx);
  return dtype == MSDA_B200_BF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(plane)[idx])
                                 : __ldg(reinterpret_cast<const float*>(plane) + idx);
}

__global__ void __launch_bounds__(kPsThreads)
point_sample_fwd_kernel(const void* const* __restrict__ planes, const msda_b200_ps_row* __restrict__ rows,
                        const float2* __restrict__ coords, float* __restrict__ out, int K, unsigned chunks) {
  // 1-D grid, row-major: the blocks of one row are scheduled together, so its plane is fetched from DRAM once and
  // then served by L2 (with the row on the fast grid axis every resident block reads a different plane: measured
  // 24.9 GB of DRAM reads and 14 % L2 hits for the 8 700-row matcher launch, profiles/r01_notes.md)
  const long long r = blockIdx.x / chunks;
  const int k = (int)(blockIdx.x % chunks) * kPsThreads + threadIdx.x;
  if (k >= K) return;
  const msda_b200_ps_row row = rows[r];
  const void* plane = planes[r];
  const float2 c = __ldg(coords + (long long)row.coord_row * K + k);
  const Corner cx = corner_setup(c.x, row.w), cy = corner_setup(c.y, row.h);
  float acc = 0.f;
  // order of accumulation as in ATen's kernel: nw, ne, sw, se
  const long long base = (long long)cy.i0 * row.w + cx.i0;
  if (cy.w0 != 0.f) {
    if (cx.w0 != 0.f) acc = fmaf(load_plane(plane, row.dtype, base), cx.w0 * cy.w0, acc);
    if (cx.w1 != 0.f) acc = fmaf(load_plane(plane, row.dtype, base + 1), cx.w1 * cy.w0, acc);
  }
  if (cy.w1 != 0.f) {
    if (cx.w0 != 0.f) acc = fmaf(load_plane(plane, row.dtype, base + row.w), cx.w0 * cy.w1, acc);
    if (cx.w1 != 0.f) acc = fmaf(load_plane(plane, row.dtype, base + row.w + 1), cx.w1 * cy.w1, acc);
  }
  out[r * K + k] = acc;
}

__global__ void __launch_bounds__(kPsThreads)
point_sample_bwd_kernel(float* const* __restrict__ grad_planes, const msda_b200_ps_row* __restrict__ rows,
                        const float2* __restrict__ coords, const float* __restrict__ grad_out, int K, unsigned chunks) {
  const long long r = blockIdx.x / chunks;
  const int k = (int)(blockIdx.x % chunks) * kPsThreads + threadIdx.x;
  if (k >= K) return;
  float* gp = grad_planes[r];
  if (gp == nullptr) return;
  const msda_b200_ps_row row = rows[r];
  const float2 c = __ldg(coords + (long long)row.coord_row * K + k);
  const float g = grad_out[r * K + k];
  const Corner cx = corner_setup(c.x, row.w), cy = corner_setup(c.y, row.h);
  const long long base = (long long)cy.i0 * row.w + cx.i0;
  if (cy.w0 != 0.f) {
    if (cx.w0 != 0.f) atomicAdd(gp + base, g * (cx.w0 * cy.w0));
    if (cx.w1 != 0.f) atomicAdd(gp + base + 1, g * (cx.w1 * cy.w0));
  }
  if (cy.w1 != 0.f) {
    if (cx.w0 != 0.f) atomicAdd(gp + base + row.w, g * (cx.w0 * cy.w1));
    if (cx.w1 != 0.f) atomicAdd(gp + base + row.w + 1, g * (cx.w1 * cy.w1));
  }
}

int check(const char* what) {
  const cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) return MSDA_B200_OK;
  msda_b200_internal_fail(MSDA_B200_ERR_CUDA, cudaGetErrorString(e));
  (void)what;
  return MSDA_B200_ERR_CUDA;
}

int check_args(const void* a, const void* b, const void* c, const void* d, int64_t R, int32_t K) {
  if (R < 0 || K < 0) return msda_b200_internal_fail(MSDA_B200_ERR_INVALID, "point_sample: negative row / point count");
  if (R > 0x7fffffffll) return msda_b200_internal_fail(MSDA_B200_ERR_UNSUPPORTED, "point_sample: more than 2^31-1 rows");
  if (R != 0 && K != 0 && (!a || !b || !c || !d))
    return msda_b200_internal_fail(MSDA_B200_ERR_INVALID, "point_sample: NULL pointer");
  return MSDA_B200_OK;
}

}  // namespace

extern "C" int msda_b200_point_sample_forward(const void* const* planes, const msda_b200_ps_row* rows, const float* coords,
                                              float* out, int64_t R, int32_t K, void* stream) {
  if (int rc = check_args(planes, rows, coords, out, R, K)) return rc;
  if (R == 0 || K == 0) return MSDA_B200_OK;
  const unsigned chunks = (unsigned)((K + kPsThreads - 1) / kPsThreads);
  if (R * (long long)chunks > 0x7fffffffll)
    return msda_b200_internal_fail(MSDA_B200_ERR_UNSUPPORTED, "point_sample: rows x points exceeds the grid limit");
  point_sample_fwd_kernel<<<(unsigned)(R * chunks), kPsThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      planes, rows, reinterpret_cast<const float2*>(coords), out, K, chunks);
  msda_b200_internal_count_launch();
  return check("point_sample_forward");
}

extern "C" int msda_b200_point_sample_backward(float* const* grad_planes, const msda_b200_ps_row* rows,
                                               const float* coords, const float* grad_out, int64_t R, int32_t K,
                                               void* stream) {
  if (int rc = check_args(grad_planes, rows, coords, grad_out, R, K)) return rc;
  if (R == 0 || K == 0) return MSDA_B200_OK;
  const unsigned chunks = (unsigned)((K + kPsThreads - 1) / kPsThreads);
  if (R * (long long)chunks > 0x7fffffffll)
    return msda_b200_internal_fail(MSDA_B200_ERR_UNSUPPORTED, "point_sample: rows x points exceeds the grid limit");
  point_sample_bwd_kernel<<<(unsigned)(R * chunks), kPsThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      grad_planes, rows, reinterpret_cast<const float2*>(coords), grad_out, K, chunks);
  msda_b200_internal_count_launch();
  return check("point_sample_backward");
}
