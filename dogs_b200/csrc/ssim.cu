// fused-SSIM forward / backward (submodules/fused-ssim/ssim.cu:187-366).
//
// The reference runs 5 (forward) / 3 (backward) separable 11-tap convolutions one after the other
// through a single scratch tile: 18 / 15 block-wide barriers per channel and a scratch re-zero
// before every convolution.  Here all statistics of a 32x32 output tile are produced in ONE
// horizontal and ONE vertical pass (3 barriers per tile), one (batch, channel) plane per
// blockIdx.z.  Tap order and the rounding points (products rounded before the tap FMA,
// `conv - mu*mu`) follow the reference, so results agree to the last bits.
//
// Register blocking (round 2): a thread produces 4 neighbouring outputs of a pass from 14 inputs held in
// registers (vector loads from 16-byte aligned rows) instead of 11 scalar shared-memory loads per tap and
// output — the first version was bound by the load/store pipe (61 % / 72 % busy, ~4x more shared-memory
// loads than FMAs need).
#include "common.cuh"

namespace dogs {

constexpr int SS_BX = 32, SS_BY = 32, SS_R = 5;
constexpr int SS_SX = SS_BX + 2 * SS_R;   // 42
constexpr int SS_SY = SS_BY + 2 * SS_R;   // 42
constexpr int SS_PAD = SS_SX + 2;         // 44: rows start on 16-byte boundaries (vector loads)
constexpr int SS_THREADS = 256;
#ifndef SS_MIN_CTAS
#define SS_MIN_CTAS 5
#endif

__constant__ float SS_G[11] = {0.001028380123898387f, 0.0075987582094967365f, 0.036000773310661316f,
                               0.10936068743467331f,  0.21300552785396576f,   0.26601171493530273f,
                               0.21300552785396576f,  0.10936068743467331f,   0.036000773310661316f,
                               0.0075987582094967365f, 0.001028380123898387f};

__device__ __forceinline__ float pix_or_zero(const float* __restrict__ plane, int y, int x, int H,
                                             int W) {
  return (x >= W || y >= H || x < 0 || y < 0) ? 0.0f : plane[(size_t)y * W + x];
}


// 14 consecutive floats of a 16-byte aligned shared-memory row
__device__ __forceinline__ void load14(const float* row, float x[14]) {
  const float4 a = *reinterpret_cast<const float4*>(row), b = *reinterpret_cast<const float4*>(row + 4),
               c = *reinterpret_cast<const float4*>(row + 8);
  const float2 d = *reinterpret_cast<const float2*>(row + 12);
  x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
  x[8] = c.x; x[9] = c.y; x[10] = c.z; x[11] = c.w; x[12] = d.x; x[13] = d.y;
}
// out[o] = sum_k G[k] * x[o + k], taps in the reference's order
__device__ __forceinline__ float4 conv4(const float x[14]) {
  float o[4];
#pragma unroll
  for (int q = 0; q < 4; q++) {
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 11; k++) acc += SS_G[k] * x[q + k];
    o[q] = acc;
  }
  return make_float4(o[0], o[1], o[2], o[3]);
}
// the same down a column of xc (14 rows starting at `col0`, row pitch SS_BX)
__device__ __forceinline__ void conv4_column(const float* col0, float o[4]) {
  float x[14];
#pragma unroll
  for (int i = 0; i < 14; i++) x[i] = col0[i * SS_BX];
#pragma unroll
  for (int q = 0; q < 4; q++) {
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 11; k++) acc += SS_G[k] * x[q + k];
    o[q] = acc;
  }
}

// Workspace of the fused photometric loss: running sums + a ticket counter.  Zero-initialised once by
// the caller; the last CTA of every launch folds the sums into the loss and zeroes it again.
struct LossWorkspace { double sum_l1, sum_ssim; unsigned int ticket; unsigned int pad; };

// LOSS = true: instead of the SSIM map the kernel produces the scalar training loss of
// gaussian_trainer.py:403-405, (1-lambda) * mean|img1-img2| + lambda * (1 - mean SSIM), in one launch.
template <bool LOSS>
__global__ void __launch_bounds__(SS_THREADS, SS_MIN_CTAS)
fssim_fwd_kernel(int H, int W, float C1, float C2, const float* __restrict__ img1,
                 const float* __restrict__ img2, float* __restrict__ ssim_map,
                 float* __restrict__ dm_dmu1, float* __restrict__ dm_dsigma1_sq,
                 float* __restrict__ dm_dsigma12, LossWorkspace* __restrict__ ws, float lambda,
                 double inv_count, float* __restrict__ loss_out) {
  __shared__ __align__(16) float p1[SS_SY][SS_PAD];
  __shared__ __align__(16) float p2[SS_SY][SS_PAD];
  __shared__ __align__(16) float xc[5][SS_SY][SS_BX];
  const int tid = threadIdx.x;
  const size_t plane = (size_t)blockIdx.z * H * W;
  const float* a = img1 + plane;
  const float* b = img2 + plane;
  const int y0 = blockIdx.y * SS_BY - SS_R, x0 = blockIdx.x * SS_BX - SS_R;
  {
    // every global load of the tile is issued before the first one is consumed (7 elements per thread and
    // image; a rolled loop paid the memory latency seven times over)
    constexpr int NLD = (SS_SY * SS_SX + SS_THREADS - 1) / SS_THREADS;
    float r1[NLD], r2[NLD];
#pragma unroll
    for (int i = 0; i < NLD; i++) {
      const int t = tid + i * SS_THREADS;
      const int ly = t / SS_SX, lx = t - ly * SS_SX;
      const bool in = t < SS_SY * SS_SX;
      r1[i] = in ? pix_or_zero(a, y0 + ly, x0 + lx, H, W) : 0.f;
      r2[i] = in ? pix_or_zero(b, y0 + ly, x0 + lx, H, W) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < NLD; i++) {
      const int t = tid + i * SS_THREADS;
      const int ly = t / SS_SX, lx = t - ly * SS_SX;
      if (t < SS_SY * SS_SX) { p1[ly][lx] = r1[i]; p2[ly][lx] = r2[i]; }
    }
  }
  __syncthreads();
  // horizontal pass: (row, 4 columns) per thread
  for (int t = tid; t < SS_SY * (SS_BX / 4); t += SS_THREADS) {
    const int ly = t >> 3, x0 = (t & 7) * 4;
    float u[14], v[14], w[14];
    load14(&p1[ly][x0], u);
    *reinterpret_cast<float4*>(&xc[0][ly][x0]) = conv4(u);
#pragma unroll
    for (int i = 0; i < 14; i++) w[i] = u[i] * u[i];
    *reinterpret_cast<float4*>(&xc[1][ly][x0]) = conv4(w);
    load14(&p2[ly][x0], v);
#pragma unroll
    for (int i = 0; i < 14; i++) w[i] = u[i] * v[i];
    *reinterpret_cast<float4*>(&xc[4][ly][x0]) = conv4(w);
    *reinterpret_cast<float4*>(&xc[2][ly][x0]) = conv4(v);
#pragma unroll
    for (int i = 0; i < 14; i++) w[i] = v[i] * v[i];
    *reinterpret_cast<float4*>(&xc[3][ly][x0]) = conv4(w);
  }
  __syncthreads();
  // vertical pass: (4 rows, column) per thread
  const int lx = tid & 31;
  const int px = blockIdx.x * SS_BX + lx;
  float acc_l1 = 0.f, acc_m = 0.f;
  const int ly0 = (tid >> 5) * 4;
  float qq[5][4];
#pragma unroll
  for (int c = 0; c < 5; c++) conv4_column(&xc[c][ly0][lx], qq[c]);
#pragma unroll
  for (int r = 0; r < 4; r++) {
    const int ly = ly0 + r;
    const int py = blockIdx.y * SS_BY + ly;
    const float q[5] = {qq[0][r], qq[1][r], qq[2][r], qq[3][r], qq[4][r]};
    const float mu1 = q[0];
    const float sigma1_sq = q[1] - mu1 * mu1;
    const float mu2 = q[2];
    const float sigma2_sq = q[3] - mu2 * mu2;
    const float sigma12 = q[4] - mu1 * mu2;
    const float mu1_sq = mu1 * mu1;
    const float mu2_sq = mu2 * mu2;
    const float mu1_mu2 = mu1 * mu2;
    const float C = (2.0f * mu1_mu2 + C1);
    const float D = (2.0f * sigma12 + C2);
    const float A = (mu1_sq + mu2_sq + C1);
    const float B = (sigma1_sq + sigma2_sq + C2);
    const float m = (C * D) / (A * B);
    if (px < W && py < H) {
      const size_t gi = plane + (size_t)py * W + px;
      if (LOSS) {
        acc_m += m;
        acc_l1 += fabsf(p1[ly + SS_R][lx + SS_R] - p2[ly + SS_R][lx + SS_R]);
      } else {
        ssim_map[gi] = m;
      }
      if (dm_dmu1) {
        dm_dmu1[gi] = ((mu2 * 2.0f * D) / (A * B) - (mu2 * 2.0f * C) / (A * B) -
                       (mu1 * 2.0f * C * D) / (A * A * B) + (mu1 * 2.0f * C * D) / (A * B * B));
        dm_dsigma1_sq[gi] = ((-C * D) / (A * B * B));
        dm_dsigma12[gi] = ((2 * C) / (A * B));
      }
    }
  }
  if (LOSS) {
    __shared__ double red[2][SS_THREADS / 32];
    double dl1 = acc_l1, dm = acc_m;   // 4 pixels per thread in float, everything above in double
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      dl1 += __shfl_xor_sync(0xffffffffu, dl1, o);
      dm += __shfl_xor_sync(0xffffffffu, dm, o);
    }
    if ((tid & 31) == 0) { red[0][tid >> 5] = dl1; red[1][tid >> 5] = dm; }
    __syncthreads();
    if (tid == 0) {
      double a = 0.0, b = 0.0;
#pragma unroll
      for (int w = 0; w < SS_THREADS / 32; w++) { a += red[0][w]; b += red[1][w]; }
      atomicAdd(&ws->sum_l1, a);
      atomicAdd(&ws->sum_ssim, b);
      __threadfence();
      const unsigned int total = gridDim.x * gridDim.y * gridDim.z;
      if (atomicAdd(&ws->ticket, 1u) == total - 1) {   // last CTA: every partial sum is visible
        __threadfence();
        const double s_l1 = *reinterpret_cast<volatile double*>(&ws->sum_l1);
        const double s_m = *reinterpret_cast<volatile double*>(&ws->sum_ssim);
        *loss_out = (float)((1.0 - (double)lambda) * (s_l1 * inv_count) + (double)lambda * (1.0 - s_m * inv_count));
        ws->sum_l1 = 0.0; ws->sum_ssim = 0.0; ws->ticket = 0u;   // ready for the next launch
      }
    }
  }
}

// LOSS = true: backward of the fused photometric loss.  dL/dmap is the constant w_ssim (= -lambda/N)
// and the L1 term w_l1 * sign(img1 - img2) (= (1-lambda)/N * sgn) is added to the same store; both are
// scaled by the upstream scalar *grad_loss when it is given.
template <bool LOSS>
__global__ void __launch_bounds__(SS_THREADS, SS_MIN_CTAS)
fssim_bwd_kernel(int H, int W, const float* __restrict__ img1, const float* __restrict__ img2,
                 const float* __restrict__ dL_dmap, const float* __restrict__ dm_dmu1,
                 const float* __restrict__ dm_dsigma1_sq, const float* __restrict__ dm_dsigma12,
                 float* __restrict__ dL_dimg1, float w_ssim, float w_l1,
                 const float* __restrict__ grad_loss) {
  if (LOSS && grad_loss) { const float g = *grad_loss; w_ssim *= g; w_l1 *= g; }
  __shared__ __align__(16) float t1[SS_SY][SS_PAD];
  __shared__ __align__(16) float t2[SS_SY][SS_PAD];
  __shared__ __align__(16) float t3[SS_SY][SS_PAD];
  __shared__ __align__(16) float xc[3][SS_SY][SS_BX];
  const int tid = threadIdx.x;
  const size_t plane = (size_t)blockIdx.z * H * W;
  const int y0 = blockIdx.y * SS_BY - SS_R, x0 = blockIdx.x * SS_BX - SS_R;
  {
    constexpr int NLD = (SS_SY * SS_SX + SS_THREADS - 1) / SS_THREADS;
    float rd[NLD], r1[NLD], r2[NLD], r3[NLD];
#pragma unroll
    for (int i = 0; i < NLD; i++) {
      const int t = tid + i * SS_THREADS;
      const int ly = t / SS_SX, lx = t - ly * SS_SX;
      const int y = y0 + ly, x = x0 + lx;
      rd[i] = r1[i] = r2[i] = r3[i] = 0.f;
      if (t < SS_SY * SS_SX && !(x >= W || y >= H || x < 0 || y < 0)) {
        const size_t gi = plane + (size_t)y * W + x;
        rd[i] = LOSS ? w_ssim : dL_dmap[gi]; r1[i] = dm_dmu1[gi]; r2[i] = dm_dsigma1_sq[gi]; r3[i] = dm_dsigma12[gi];
      }
    }
#pragma unroll
    for (int i = 0; i < NLD; i++) {
      const int t = tid + i * SS_THREADS;
      const int ly = t / SS_SX, lx = t - ly * SS_SX;
      if (t < SS_SY * SS_SX) {
        t1[ly][lx] = r1[i] * rd[i];
        t2[ly][lx] = r2[i] * rd[i];
        t3[ly][lx] = r3[i] * rd[i];
      }
    }
  }
  __syncthreads();
  for (int t = tid; t < SS_SY * (SS_BX / 4); t += SS_THREADS) {
    const int ly = t >> 3, x0 = (t & 7) * 4;
    float x[14];
    load14(&t1[ly][x0], x);
    *reinterpret_cast<float4*>(&xc[0][ly][x0]) = conv4(x);
    load14(&t2[ly][x0], x);
    *reinterpret_cast<float4*>(&xc[1][ly][x0]) = conv4(x);
    load14(&t3[ly][x0], x);
    *reinterpret_cast<float4*>(&xc[2][ly][x0]) = conv4(x);
  }
  __syncthreads();
  const int lx = tid & 31;
  const int px = blockIdx.x * SS_BX + lx;
  const int ly0 = (tid >> 5) * 4;
  float qq[3][4];
#pragma unroll
  for (int c = 0; c < 3; c++) conv4_column(&xc[c][ly0][lx], qq[c]);
#pragma unroll
  for (int r = 0; r < 4; r++) {
    const int py = blockIdx.y * SS_BY + ly0 + r;
    const float q[3] = {qq[0][r], qq[1][r], qq[2][r]};
    if (px < W && py < H) {
      const size_t gi = plane + (size_t)py * W + px;
      const float pix1 = img1[gi], pix2 = img2[gi];
      float dL_dpix = 0.0f;
      dL_dpix += q[0];
      dL_dpix += pix1 * 2.0f * q[1];
      dL_dpix += pix2 * q[2];
      if (LOSS) {   // + (1-lambda)/N * sgn(img1 - img2), as autograd sums the two branches
        const float d = pix1 - pix2;
        dL_dpix += d > 0.f ? w_l1 : (d < 0.f ? -w_l1 : 0.f);
      }
      dL_dimg1[gi] = dL_dpix;
    }
  }
}

}  // namespace dogs

using namespace dogs;

extern "C" int fssim_forward(int32_t B, int32_t CH, int32_t H, int32_t W, float C1, float C2,
                             const float* img1, const float* img2, float* ssim_map, float* dm_dmu1,
                             float* dm_dsigma1_sq, float* dm_dsigma12, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (B <= 0 || CH <= 0 || H <= 0 || W <= 0) return 0;
  if (!img1 || !img2 || !ssim_map) { set_error("fssim_forward: null pointer"); return 1; }
  if ((dm_dmu1 == nullptr) != (dm_dsigma1_sq == nullptr) || (dm_dmu1 == nullptr) != (dm_dsigma12 == nullptr)) {
    set_error("fssim_forward: pass all three derivative maps or none");
    return 1;
  }
  if ((long long)B * CH > 65535) { set_error("fssim_forward: B*CH > 65535"); return 1; }
  dim3 grid((W + SS_BX - 1) / SS_BX, (H + SS_BY - 1) / SS_BY, B * CH);
  fssim_fwd_kernel<false><<<grid, SS_THREADS, 0, s>>>(H, W, C1, C2, img1, img2, ssim_map, dm_dmu1,
                                                      dm_dsigma1_sq, dm_dsigma12, nullptr, 0.f, 0.0, nullptr);
  return check_launch("fssim_forward", false, s);
}

extern "C" int fssim_backward(int32_t B, int32_t CH, int32_t H, int32_t W, float C1, float C2,
                              const float* img1, const float* img2, const float* dL_dmap,
                              const float* dm_dmu1, const float* dm_dsigma1_sq,
                              const float* dm_dsigma12, float* dL_dimg1, void* stream) {
  (void)C1; (void)C2;
  cudaStream_t s = (cudaStream_t)stream;
  if (B <= 0 || CH <= 0 || H <= 0 || W <= 0) return 0;
  if (!img1 || !img2 || !dL_dmap || !dm_dmu1 || !dm_dsigma1_sq || !dm_dsigma12 || !dL_dimg1) {
    set_error("fssim_backward: null pointer");
    return 1;
  }
  if ((long long)B * CH > 65535) { set_error("fssim_backward: B*CH > 65535"); return 1; }
  dim3 grid((W + SS_BX - 1) / SS_BX, (H + SS_BY - 1) / SS_BY, B * CH);
  fssim_bwd_kernel<false><<<grid, SS_THREADS, 0, s>>>(H, W, img1, img2, dL_dmap, dm_dmu1, dm_dsigma1_sq,
                                                      dm_dsigma12, dL_dimg1, 0.f, 0.f, nullptr);
  return check_launch("fssim_backward", false, s);
}

extern "C" int fssim_l1_loss_forward(int32_t B, int32_t CH, int32_t H, int32_t W, float C1, float C2,
                                     float lambda_dssim, const float* img1, const float* img2,
                                     float* dm_dmu1, float* dm_dsigma1_sq, float* dm_dsigma12,
                                     void* workspace, float* loss_out, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (B <= 0 || CH <= 0 || H <= 0 || W <= 0) { set_error("fssim_l1_loss_forward: empty image"); return 1; }
  if (!img1 || !img2 || !workspace || !loss_out) { set_error("fssim_l1_loss_forward: null pointer"); return 1; }
  if ((dm_dmu1 == nullptr) != (dm_dsigma1_sq == nullptr) || (dm_dmu1 == nullptr) != (dm_dsigma12 == nullptr)) {
    set_error("fssim_l1_loss_forward: pass all three derivative maps or none");
    return 1;
  }
  if ((long long)B * CH > 65535) { set_error("fssim_l1_loss_forward: B*CH > 65535"); return 1; }
  dim3 grid((W + SS_BX - 1) / SS_BX, (H + SS_BY - 1) / SS_BY, B * CH);
  const double inv_count = 1.0 / ((double)B * CH * H * W);
  fssim_fwd_kernel<true><<<grid, SS_THREADS, 0, s>>>(H, W, C1, C2, img1, img2, nullptr, dm_dmu1, dm_dsigma1_sq,
                                                     dm_dsigma12, (LossWorkspace*)workspace, lambda_dssim,
                                                     inv_count, loss_out);
  return check_launch("fssim_l1_loss_forward", false, s);
}

extern "C" int fssim_l1_loss_backward(int32_t B, int32_t CH, int32_t H, int32_t W, float w_ssim, float w_l1,
                                      const float* img1, const float* img2, const float* dm_dmu1,
                                      const float* dm_dsigma1_sq, const float* dm_dsigma12,
                                      const float* grad_loss, float* dL_dimg1, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (B <= 0 || CH <= 0 || H <= 0 || W <= 0) return 0;
  if (!img1 || !img2 || !dm_dmu1 || !dm_dsigma1_sq || !dm_dsigma12 || !dL_dimg1) {
    set_error("fssim_l1_loss_backward: null pointer");
    return 1;
  }
  if ((long long)B * CH > 65535) { set_error("fssim_l1_loss_backward: B*CH > 65535"); return 1; }
  dim3 grid((W + SS_BX - 1) / SS_BX, (H + SS_BY - 1) / SS_BY, B * CH);
  fssim_bwd_kernel<true><<<grid, SS_THREADS, 0, s>>>(H, W, img1, img2, nullptr, dm_dmu1, dm_dsigma1_sq,
                                                     dm_dsigma12, dL_dimg1, w_ssim, w_l1, grad_loss);
  return check_launch("fssim_l1_loss_backward", false, s);
}
