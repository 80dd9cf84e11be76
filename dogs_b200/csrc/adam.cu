// Sparse Adam step (SparseGaussianAdam.step -> adamUpdateCUDA, cuda_rasterizer/adam.cu:9-38):
// only Gaussians with visible[g] are touched; beta powers are not bias-corrected (as in the
// reference).  HBM-bound: 28 B per updated element.  Each thread handles 4 consecutive elements
// with 16-byte loads/stores when the tensors allow it; the update expression is kept in the
// reference's shape so the result is bit-identical.
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

}  // namespace dogs

using namespace dogs;

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
