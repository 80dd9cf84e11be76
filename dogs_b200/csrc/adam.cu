// Sparse Adam step (SparseGaussianAdam.step -> adamUpdateCUDA, cuda_rasterizer/adam.cu:9-38):
// only Gaussians with visible[g] are touched; beta powers are not bias-corrected (as in the
// reference).  HBM-bound: 28 B per updated element.  Each thread handles 4 consecutive elements
// with 16-byte loads/stores when the tensors allow it; the update expression is kept in the
// reference's shape so the result is bit-identical.
#include <string.h>
#include "common.cuh"

namespace dogs {

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, float lr, float b1,
                                         float b2, float eps) {
  // Explicit roundings = the FMA placement nvcc chose for the reference kernel (adam.cu:29-30 SASS):
  // fma(g, 1-b1, b1*m) and fma(g, (1-b2)*g, b2*v); spelled out so inlining cannot change them.
  m = __fmaf_rn(g, 1.0f - b1, __fmul_rn(b1, m));
  v = __fmaf_rn(g, __fmul_rn(1.0f - b2, g), __fmul_rn(b2, v));
  float step = -lr * m / (sqrt(v) + eps);
  p += step;
}

__global__ void __launch_bounds__(256)
sparse_adam_vec4_kernel(float4* __restrict__ param, const float4* __restrict__ grad,
                        float4* __restrict__ exp_avg, float4* __restrict__ exp_avg_sq,
                        const uint8_t* __restrict__ visible, float lr, float b1, float b2, float eps,
                        uint32_t N, uint32_t M, uint32_t n4) {
  uint32_t i4 = blockIdx.x * blockDim.x + threadIdx.x;
  if (i4 >= n4) return;
  const uint32_t e0 = i4 * 4;
  const uint32_t g0 = e0 / M, g3 = (e0 + 3) / M;
  bool vis[4];
  bool any = false;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    uint32_t g = (g0 == g3) ? g0 : (e0 + k) / M;
    vis[k] = g < N && visible[g];
    any |= vis[k];
  }
  if (!any) return;
  float4 p = param[i4], gr = grad[i4], m = exp_avg[i4], v = exp_avg_sq[i4];
  if (vis[0]) adam_one(p.x, gr.x, m.x, v.x, lr, b1, b2, eps);
  if (vis[1]) adam_one(p.y, gr.y, m.y, v.y, lr, b1, b2, eps);
  if (vis[2]) adam_one(p.z, gr.z, m.z, v.z, lr, b1, b2, eps);
  if (vis[3]) adam_one(p.w, gr.w, m.w, v.w, lr, b1, b2, eps);
  param[i4] = p;
  exp_avg[i4] = m;
  exp_avg_sq[i4] = v;
}

__global__ void __launch_bounds__(256)
sparse_adam_scalar_kernel(float* __restrict__ param, const float* __restrict__ grad,
                          float* __restrict__ exp_avg, float* __restrict__ exp_avg_sq,
                          const uint8_t* __restrict__ visible, float lr, float b1, float b2,
                          float eps, uint32_t N, uint32_t M, uint32_t first, uint32_t total) {
  uint32_t i = first + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  uint32_t g = i / M;
  if (g >= N || !visible[g]) return;
  float p = param[i], m = exp_avg[i], v = exp_avg_sq[i];
  adam_one(p, grad[i], m, v, lr, b1, b2, eps);
  param[i] = p;
  exp_avg[i] = m;
  exp_avg_sq[i] = v;
}

// All parameter tensors of the model in ONE launch: the reference's optimizer step is six
// adamUpdate launches per iteration (SparseGaussianAdam.step loops over its param groups, __init__.py:303-332),
// each reading the visibility mask again.  Work is laid out as 16-byte chunks over the concatenation of the
// tensors; a chunk never straddles two tensors, the tail of a tensor whose length is not a multiple of 4 is
// handled element-wise.  Visibility comes either from a bool mask or straight from the rasterizer's radii
// (visible = radii > 0, gaussian_trainer.py:474-475), which saves the comparison kernel as well.
constexpr int ADAM_MAX_TENSORS = 8;
struct AdamMultiArgs {
  float* param[ADAM_MAX_TENSORS];
  const float* grad[ADAM_MAX_TENSORS];
  float* exp_avg[ADAM_MAX_TENSORS];
  float* exp_avg_sq[ADAM_MAX_TENSORS];
  uint32_t M[ADAM_MAX_TENSORS];
  float lr[ADAM_MAX_TENSORS];
  float eps[ADAM_MAX_TENSORS];
  unsigned long long chunk_end[ADAM_MAX_TENSORS];   // exclusive prefix ends, in 4-float chunks
  int n;
};

__global__ void __launch_bounds__(256)
sparse_adam_multi_kernel(const __grid_constant__ AdamMultiArgs a, const uint8_t* __restrict__ visible,
                         const int32_t* __restrict__ radii, float b1, float b2, uint32_t N) {
  const unsigned long long c = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= a.chunk_end[a.n - 1]) return;
  int t = 0;
#pragma unroll
  for (int q = 0; q < ADAM_MAX_TENSORS - 1; q++) t += (q < a.n - 1 && c >= a.chunk_end[q]) ? 1 : 0;
  const unsigned long long lc = c - (t ? a.chunk_end[t - 1] : 0ull);
  const uint32_t M = a.M[t];
  const uint32_t len = N * M;                       // < 2^32 per tensor (checked by the host)
  const uint32_t e0 = (uint32_t)lc * 4u;
  const uint32_t cnt = min(4u, len - e0);
  const uint32_t g0 = e0 / M, g3 = (e0 + cnt - 1) / M;
  bool vis[4] = {false, false, false, false};
  bool any = false;
  for (uint32_t k = 0; k < cnt; k++) {
    const uint32_t g = (g0 == g3) ? g0 : (e0 + k) / M;
    vis[k] = radii ? radii[g] > 0 : visible[g] != 0;
    any |= vis[k];
  }
  if (!any) return;
  const float lr = a.lr[t], eps = a.eps[t];
  float* __restrict__ P = a.param[t];
  const float* __restrict__ G = a.grad[t];
  float* __restrict__ Mo = a.exp_avg[t];
  float* __restrict__ Vo = a.exp_avg_sq[t];
  if (cnt == 4) {
    float4 p = *reinterpret_cast<float4*>(P + e0), m = *reinterpret_cast<float4*>(Mo + e0), v = *reinterpret_cast<float4*>(Vo + e0);
    const float4 gr = *reinterpret_cast<const float4*>(G + e0);
    if (vis[0]) adam_one(p.x, gr.x, m.x, v.x, lr, b1, b2, eps);
    if (vis[1]) adam_one(p.y, gr.y, m.y, v.y, lr, b1, b2, eps);
    if (vis[2]) adam_one(p.z, gr.z, m.z, v.z, lr, b1, b2, eps);
    if (vis[3]) adam_one(p.w, gr.w, m.w, v.w, lr, b1, b2, eps);
    *reinterpret_cast<float4*>(P + e0) = p;
    *reinterpret_cast<float4*>(Mo + e0) = m;
    *reinterpret_cast<float4*>(Vo + e0) = v;
  } else {
    for (uint32_t k = 0; k < cnt; k++) {
      if (!vis[k]) continue;
      float p = P[e0 + k], m = Mo[e0 + k], v = Vo[e0 + k];
      adam_one(p, G[e0 + k], m, v, lr, b1, b2, eps);
      P[e0 + k] = p; Mo[e0 + k] = m; Vo[e0 + k] = v;
    }
  }
}

}  // namespace dogs

using namespace dogs;

extern "C" int dgr_sparse_adam_multi(const DgrAdamTensor* tensors, int32_t n_tensors, const uint8_t* visible,
                                     const int32_t* radii, float b1, float b2, uint32_t N, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (n_tensors <= 0 || N == 0) return 0;
  if (n_tensors > ADAM_MAX_TENSORS) { set_error("sparse_adam_multi: at most %d tensors per launch", ADAM_MAX_TENSORS); return 1; }
  if (!tensors || (!visible && !radii)) { set_error("sparse_adam_multi: null pointer"); return 1; }
  AdamMultiArgs a;
  memset(&a, 0, sizeof(a));
  a.n = n_tensors;
  unsigned long long end = 0;
  for (int t = 0; t < n_tensors; t++) {
    const DgrAdamTensor& x = tensors[t];
    if (!x.param || !x.grad || !x.exp_avg || !x.exp_avg_sq || x.M == 0) { set_error("sparse_adam_multi: tensor %d is incomplete", t); return 1; }
    if ((((uintptr_t)x.param) | ((uintptr_t)x.grad) | ((uintptr_t)x.exp_avg) | ((uintptr_t)x.exp_avg_sq)) & 15) {
      set_error("sparse_adam_multi: tensor %d is not 16-byte aligned", t);
      return 1;
    }
    a.param[t] = x.param; a.grad[t] = x.grad; a.exp_avg[t] = x.exp_avg; a.exp_avg_sq[t] = x.exp_avg_sq;
    a.M[t] = x.M; a.lr[t] = x.lr; a.eps[t] = x.eps;
    if ((unsigned long long)N * x.M > 0xffffffffull) { set_error("sparse_adam_multi: tensor %d too large", t); return 1; }
    end += ((unsigned long long)N * x.M + 3) / 4;
    a.chunk_end[t] = end;
  }
  for (int t = n_tensors; t < ADAM_MAX_TENSORS; t++) a.chunk_end[t] = end;
  if (end > 0x7fffffffull * 256) { set_error("sparse_adam_multi: too many elements"); return 1; }
  sparse_adam_multi_kernel<<<(unsigned)((end + 255) / 256), 256, 0, s>>>(a, visible, radii, b1, b2, N);
  return check_launch("sparse_adam_multi", false, s);
}

extern "C" int dgr_sparse_adam(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                               const uint8_t* visible, float lr, float b1, float b2, float eps,
                               uint32_t N, uint32_t M, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  const uint64_t total64 = (uint64_t)N * M;
  if (total64 == 0) return 0;
  if (total64 > 0xffffffffull) { set_error("sparse_adam: tensor too large"); return 1; }
  if (!param || !grad || !exp_avg || !exp_avg_sq || !visible) { set_error("sparse_adam: null pointer"); return 1; }
  const uint32_t total = (uint32_t)total64;
  const bool aligned = ((((uintptr_t)param) | ((uintptr_t)grad) | ((uintptr_t)exp_avg) | ((uintptr_t)exp_avg_sq)) & 15) == 0;
  uint32_t n4 = aligned ? total / 4 : 0;
  if (n4)
    sparse_adam_vec4_kernel<<<(n4 + 255) / 256, 256, 0, s>>>(
        (float4*)param, (const float4*)grad, (float4*)exp_avg, (float4*)exp_avg_sq, visible, lr, b1,
        b2, eps, N, M, n4);
  const uint32_t rest = total - n4 * 4;
  if (rest)
    sparse_adam_scalar_kernel<<<(rest + 255) / 256, 256, 0, s>>>(param, grad, exp_avg, exp_avg_sq,
                                                                visible, lr, b1, b2, eps, N, M,
                                                                n4 * 4, total);
  return check_launch("sparse_adam", false, s);
}
