// gemm_emul.cu -- feasibility probe (not part of the product): fp32 GEMMs of the encoder layer's shapes through cuBLASLt
// with CUBLAS_COMPUTE_32F (what torch runs for fp32 tensors), CUBLAS_COMPUTE_32F_FAST_TF32 and
// CUBLAS_COMPUTE_32F_EMULATED_16BFX9 (CUDA 12.9: fp32 emulated on the bf16 tensor cores, 9 products) -- time and error
// against an fp64 reference on sampled outputs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gemm_emul gemm_emul.cu -lcublasLt && ./gemm_emul
#include <cublasLt.h>
#include <cuda_runtime.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
#define CB(x) do { cublasStatus_t s = (x); if (s != CUBLAS_STATUS_SUCCESS) { printf("cuBLAS error %d at %d\n", (int)s, __LINE__); return -1.f; } } while (0)

static cublasLtHandle_t lt;
static void* ws; static size_t ws_bytes = 256u << 20;

// y[M,N] = x[M,K] @ w[N,K]^T, all row-major (column-major: C[N,M] = op_T(w)[N,K] * x^T[K,M])
static float run(cublasComputeType_t ct, int M, int N, int K, const float* x, const float* w, float* y, int iters) {
  cublasLtMatmulDesc_t op; cublasLtMatrixLayout_t la, lb, lc; cublasLtMatmulPreference_t pref;
  CB(cublasLtMatmulDescCreate(&op, ct, CUDA_R_32F));
  cublasOperation_t T = CUBLAS_OP_T, Nn = CUBLAS_OP_N;
  CB(cublasLtMatmulDescSetAttribute(op, CUBLASLT_MATMUL_DESC_TRANSA, &T, sizeof(T)));
  CB(cublasLtMatmulDescSetAttribute(op, CUBLASLT_MATMUL_DESC_TRANSB, &Nn, sizeof(Nn)));
  CB(cublasLtMatrixLayoutCreate(&la, CUDA_R_32F, K, N, K));
  CB(cublasLtMatrixLayoutCreate(&lb, CUDA_R_32F, K, M, K));
  CB(cublasLtMatrixLayoutCreate(&lc, CUDA_R_32F, N, M, N));
  CB(cublasLtMatmulPreferenceCreate(&pref));
  CB(cublasLtMatmulPreferenceSetAttribute(pref, CUBLASLT_MATMUL_PREF_MAX_WORKSPACE_BYTES, &ws_bytes, sizeof(ws_bytes)));
  cublasLtMatmulHeuristicResult_t h; int found = 0;
  CB(cublasLtMatmulAlgoGetHeuristic(lt, op, la, lb, lc, lc, pref, 1, &h, &found));
  if (!found) { printf("no algorithm\n"); return -1.f; }
  const float one = 1.f, zero = 0.f;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) CB(cublasLtMatmul(lt, op, &one, w, la, x, lb, &zero, y, lc, y, lc, &h.algo, ws, ws_bytes, 0));
  cudaEventRecord(e0);
  for (int i = 0; i < iters; ++i) CB(cublasLtMatmul(lt, op, &one, w, la, x, lb, &zero, y, lc, y, lc, &h.algo, ws, ws_bytes, 0));
  cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  cublasLtMatmulPreferenceDestroy(pref); cublasLtMatrixLayoutDestroy(la); cublasLtMatrixLayoutDestroy(lb);
  cublasLtMatrixLayoutDestroy(lc); cublasLtMatmulDescDestroy(op);
  return ms / iters;
}

int main() {
  if (cublasLtCreate(&lt) != CUBLAS_STATUS_SUCCESS) { printf("cublasLtCreate failed\n"); return 1; }
  CK(cudaMalloc(&ws, ws_bytes));
  printf("cublasLt version %zu\n", cublasLtGetVersion());
  const int M = 172032;
  const int shapes[][2] = {{256, 256}, {1024, 256}, {256, 1024}, {288, 256}};  // (N, K): value/out proj, fc1, fc2, offsets+weights
  for (auto& s : shapes) {
    const int N = s[0], K = s[1];
    std::vector<float> hx((size_t)M * K), hw((size_t)N * K);
    unsigned r = 12345u;
    auto rnd = [&]() { r = r * 1664525u + 1013904223u; return ((r >> 8) & 0xffff) / 65536.f - 0.5f; };
    for (auto& v : hx) v = rnd() * 4.f;
    for (auto& v : hw) v = rnd() * 0.25f;
    float *x, *w, *y; CK(cudaMalloc(&x, hx.size() * 4)); CK(cudaMalloc(&w, hw.size() * 4)); CK(cudaMalloc(&y, (size_t)M * N * 4));
    CK(cudaMemcpy(x, hx.data(), hx.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(w, hw.data(), hw.size() * 4, cudaMemcpyHostToDevice));
    const cublasComputeType_t cts[] = {CUBLAS_COMPUTE_32F, CUBLAS_COMPUTE_32F_FAST_TF32, CUBLAS_COMPUTE_32F_EMULATED_16BFX9};
    const char* names[] = {"32F", "32F_FAST_TF32", "32F_EMULATED_16BFX9"};
    for (int c = 0; c < 3; ++c) {
      const float ms = run(cts[c], M, N, K, x, w, y, 20);
      if (ms < 0) { printf("M=%d N=%d K=%d %-20s unsupported\n", M, N, K, names[c]); continue; }
      // error on 64 sampled rows against fp64
      std::vector<float> hy((size_t)64 * N);
      double emax = 0, ymax = 0;
      for (int i = 0; i < 64; ++i) {
        const size_t row = (size_t)i * 2687 % M;
        CK(cudaMemcpy(hy.data() + (size_t)i * N, y + row * N, (size_t)N * 4, cudaMemcpyDeviceToHost));
        for (int n = 0; n < N; ++n) {
          double acc = 0;
          for (int k = 0; k < K; ++k) acc += (double)hx[row * K + k] * (double)hw[(size_t)n * K + k];
          emax = fmax(emax, fabs(acc - hy[(size_t)i * N + n])); ymax = fmax(ymax, fabs(acc));
        }
      }
      printf("M=%d N=%4d K=%4d %-20s %7.3f ms  %7.1f TFLOP/s  max err / max |y| = %.2e\n", M, N, K, names[c], ms,
             2.0 * M * N * K / ms / 1e9, emax / ymax);
    }
    cudaFree(x); cudaFree(w); cudaFree(y);
  }
  return 0;
}
