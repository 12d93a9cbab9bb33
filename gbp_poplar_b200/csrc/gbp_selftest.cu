// Device-side unit checks of the math helpers (gbp_math.cuh) behind three C-ABI entry points: the SAME __device__
// functions the sweep kernels inline are run on caller-supplied inputs, one thread per item, so that the parity
// suite can pin each of them to the reference's helpers (ba/matlib.cpp:143-222 inv3x3 / inv6x6,
// ba/bafuncs.cpp:32-213 so3exp / hfunc / Jac) on the golden vectors of tests/golden/golden_helpers.npz -- not only
// transitively through whole sweeps.
#include <cuda_runtime.h>

#include <string>

#include "../../include/gbp_cuda.h"
#include "gbp_math.cuh"

void gbp_set_error(const std::string& s);  // host_error.cpp

namespace {

__global__ void k_test_inv6(const float* __restrict__ A, float* __restrict__ out, const int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float L[21], Ai[36];
  for (int r = 0; r < 6; ++r)
    for (int c = 0; c <= r; ++c) L[gbp::lt(r, c)] = A[(size_t)i * 36 + r * 6 + c];  // inv6x6 reads the lower triangle only
  gbp::inv6(L, Ai);
  for (int k = 0; k < 36; ++k) out[(size_t)i * 36 + k] = Ai[k];
}

__global__ void k_test_inv3(const float* __restrict__ A, float* __restrict__ out, const int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float M[9], R[9];
  for (int k = 0; k < 9; ++k) M[k] = A[(size_t)i * 9 + k];
  gbp::inv3(M, R);
  for (int k = 0; k < 9; ++k) out[(size_t)i * 9 + k] = R[k];
}

__global__ void k_test_project(const float* __restrict__ X, const float* __restrict__ P, const float4 K, float* __restrict__ hx,
                               float* __restrict__ Jkf, float* __restrict__ Jlmk, const int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float x_kf[6], x_l[3];
  for (int k = 0; k < 6; ++k) x_kf[k] = X[(size_t)i * 6 + k];
  for (int k = 0; k < 3; ++k) x_l[k] = P[(size_t)i * 3 + k];
  const float w[3] = {x_kf[3], x_kf[4], x_kf[5]};
  float R[9], num[9], den;
  gbp::cam_lin_consts(w, R, num, den);  // so3exp + the camera-only part of Jac
  const float Kf[4] = {K.x, K.y, K.z, K.w};
  float Jk[12], Jl[6], h0, h1;
  gbp::project_jac(Kf, x_kf, x_l, R, num, den, Jk, Jl, h0, h1);
  hx[(size_t)i * 2] = h0;
  hx[(size_t)i * 2 + 1] = h1;
  for (int k = 0; k < 12; ++k) Jkf[(size_t)i * 12 + k] = Jk[k];
  for (int k = 0; k < 6; ++k) Jlmk[(size_t)i * 6 + k] = Jl[k];
}

struct DevBuf {
  float* p = nullptr;
  ~DevBuf() { cudaFree(p); }
  bool alloc(size_t n) { return cudaMalloc((void**)&p, std::max<size_t>(n, 1) * sizeof(float)) == cudaSuccess; }
};

int finish(const char* what) {
  const cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    gbp_set_error(std::string(what) + ": " + cudaGetErrorString(e));
    return GBP_ERR_CUDA;
  }
  return GBP_OK;
}

int run_unary(const float* A, float* out, int n, int width, void (*kernel)(const float*, float*, int), const char* what) {
  if (!A || !out || n < 0) return GBP_ERR_ARG;
  DevBuf a, o;
  if (!a.alloc((size_t)n * width) || !o.alloc((size_t)n * width)) {
    gbp_set_error("no CUDA device available (the GBP hot path has no CPU fallback)");
    return GBP_ERR_CUDA;
  }
  cudaMemcpy(a.p, A, (size_t)n * width * sizeof(float), cudaMemcpyHostToDevice);
  if (n) kernel<<<(n + 63) / 64, 64>>>(a.p, o.p, n);
  int rc = finish(what);
  if (!rc) cudaMemcpy(out, o.p, (size_t)n * width * sizeof(float), cudaMemcpyDeviceToHost);
  return rc;
}

}  // namespace

extern "C" {

int gbp_cuda_test_inv6x6(const float* A, float* out, int n) { return run_unary(A, out, n, 36, k_test_inv6, "inv6x6 self-test"); }
int gbp_cuda_test_inv3x3(const float* A, float* out, int n) { return run_unary(A, out, n, 9, k_test_inv3, "inv3x3 self-test"); }

int gbp_cuda_test_project(const float* X, const float* P, const float* K9, float* hx, float* Jkf, float* Jlmk, int n) {
  if (!X || !P || !K9 || !hx || !Jkf || !Jlmk || n < 0) return GBP_ERR_ARG;
  DevBuf x, p, h, jk, jl;
  if (!x.alloc((size_t)n * 6) || !p.alloc((size_t)n * 3) || !h.alloc((size_t)n * 2) || !jk.alloc((size_t)n * 12) || !jl.alloc((size_t)n * 6)) {
    gbp_set_error("no CUDA device available (the GBP hot path has no CPU fallback)");
    return GBP_ERR_CUDA;
  }
  cudaMemcpy(x.p, X, (size_t)n * 6 * sizeof(float), cudaMemcpyHostToDevice);
  cudaMemcpy(p.p, P, (size_t)n * 3 * sizeof(float), cudaMemcpyHostToDevice);
  if (n) k_test_project<<<(n + 63) / 64, 64>>>(x.p, p.p, make_float4(K9[0], K9[4], K9[2], K9[5]), h.p, jk.p, jl.p, n);
  int rc = finish("projection self-test");
  if (!rc) {
    cudaMemcpy(hx, h.p, (size_t)n * 2 * sizeof(float), cudaMemcpyDeviceToHost);
    cudaMemcpy(Jkf, jk.p, (size_t)n * 12 * sizeof(float), cudaMemcpyDeviceToHost);
    cudaMemcpy(Jlmk, jl.p, (size_t)n * 6 * sizeof(float), cudaMemcpyDeviceToHost);
  }
  return rc;
}

}  // extern "C"
