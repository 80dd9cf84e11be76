// Backward rasterizer kernels (sm_100a).
//
// blend_bwd_kernel replaces PerGaussianRenderCUDA (cuda_rasterizer/backward.cu:453-658).  The
// reference walks each 32-splat bucket over the tile's 256 pixels and needs a (T, colour, depth)
// snapshot per pixel per bucket from the forward pass (5 KB / bucket, ~2.7 GB at 3 M Gaussians).
// Here every pixel thread replays its own front-to-back blend (T evolves exactly as in the
// forward pass, the entry count is bounded by n_contrib), so no snapshot is stored at all.  The
// 10 per-pair gradient terms are reduced across the warp with a transposed butterfly
// (12 shuffles instead of 50), accumulated per tile in shared memory, and flushed with three
// vector reductions (red.global.add.v4/v2.f32) per tile instance.
//
// preprocess_bwd_kernel fuses computeCov2DCUDA (backward.cu:149-326) and preprocessCUDA
// (backward.cu:398-451) and writes every output gradient exactly once (zeros for culled
// Gaussians), which removes the 308 B/Gaussian of torch::zeros memsets
// (rasterize_points.cu:196-207).
#include <string.h>
#include "common.cuh"

namespace dogs {

constexpr int ACC_STRIDE = 12;  // floats per Gaussian in the global accumulator (3 x float4)
// accumulator slots: 0 mean2D.x, 1 mean2D.y, 2 conic.x, 3 conic.y, 4 conic.w, 5 opacity,
//                    6..8 colour rgb, 9 inverse depth, 10..11 unused

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}
__device__ __forceinline__ void red_add_v2(float* p, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}

// Sum 10 per-lane values over the warp.  After the call lane l holds the total of value
// slot(l) (valid lanes only); every total lives in exactly two lanes (bit 0 of the lane id).
__device__ __forceinline__ float warp_reduce10(const float v[10], int lane) {
  const unsigned full = 0xffffffffu;
  bool hi = lane & 16;
  float k[5];
#pragma unroll
  for (int i = 0; i < 5; i++) {
    float keep = hi ? v[5 + i] : v[i];
    float send = hi ? v[i] : v[5 + i];
    k[i] = keep + __shfl_xor_sync(full, send, 16);
  }
  hi = lane & 8;
  float a0 = (hi ? k[3] : k[0]) + __shfl_xor_sync(full, hi ? k[0] : k[3], 8);
  float a1 = (hi ? k[4] : k[1]) + __shfl_xor_sync(full, hi ? k[1] : k[4], 8);
  float a2 = (hi ? 0.f : k[2]) + __shfl_xor_sync(full, hi ? k[2] : 0.f, 8);
  hi = lane & 4;
  float b0 = (hi ? a2 : a0) + __shfl_xor_sync(full, hi ? a0 : a2, 4);
  float b1 = (hi ? 0.f : a1) + __shfl_xor_sync(full, hi ? a1 : 0.f, 4);
  hi = lane & 2;
  float c = (hi ? b1 : b0) + __shfl_xor_sync(full, hi ? b0 : b1, 2);
  c += __shfl_xor_sync(full, c, 1);
  return c;
}
// Which of the 10 values lane l ends up holding (-1: none / duplicate).
__device__ __forceinline__ int warp_reduce10_slot(int lane) {
  const int b4 = (lane >> 4) & 1, b3 = (lane >> 3) & 1, b2 = (lane >> 2) & 1, b1 = (lane >> 1) & 1;
  if (lane & 1) return -1;
  const int i2 = 2 * b2 + b1;       // index among the 3 values of stage 3
  if (i2 > 2) return -1;
  const int i1 = 3 * b3 + i2;       // index among the 5 values of stage 2
  if (b3 && i2 > 1) return -1;
  return 5 * b4 + i1;
}

// ---- backward.cu:23-144 ---------------------------------------------------------------
// sh_row: this Gaussian's staged SH "rest" row (3*M floats in shared memory); it is overwritten
// with dL_dsh so the CTA can store the whole slice with coalesced 16-byte writes.
__device__ __forceinline__ V3 sh_backward(int deg, int M, V3 pos, V3 campos, uint32_t clamped,
                                          V3 dL_dcolor, float* sh_row, V3& dL_ddc) {
  V3 dir_orig = pos - campos;
  V3 dir = dir_orig / sqrtf(dot(dir_orig, dir_orig));
  V3 dL_dRGB = dL_dcolor;
  dL_dRGB.x *= (clamped & 1u) ? 0 : 1;
  dL_dRGB.y *= (clamped & 2u) ? 0 : 1;
  dL_dRGB.z *= (clamped & 4u) ? 0 : 1;
  V3 dRGBdx = {0, 0, 0}, dRGBdy = {0, 0, 0}, dRGBdz = {0, 0, 0};
  const float x = dir.x, y = dir.y, z = dir.z;
  dL_ddc = SH_C0 * dL_dRGB;
#define SHV(i) (V3{sh_row[3 * (i)], sh_row[3 * (i) + 1], sh_row[3 * (i) + 2]})
#define SHW(i, val) { V3 t_ = (val); sh_row[3 * (i)] = t_.x; sh_row[3 * (i) + 1] = t_.y; sh_row[3 * (i) + 2] = t_.z; }
  int written = 0;
  if (deg > 0) {
    const V3 s0 = SHV(0), s1 = SHV(1), s2 = SHV(2);
    dRGBdx = (-SH_C1) * s2;
    dRGBdy = (-SH_C1) * s0;
    dRGBdz = SH_C1 * s1;
    V3 s3, s4, s5, s6, s7, s8, s9, s10, s11, s12, s13, s14;
    if (deg > 1) { s3 = SHV(3); s4 = SHV(4); s5 = SHV(5); s6 = SHV(6); s7 = SHV(7); }
    if (deg > 2) { s8 = SHV(8); s9 = SHV(9); s10 = SHV(10); s11 = SHV(11); s12 = SHV(12); s13 = SHV(13); s14 = SHV(14); }
    SHW(0, (-SH_C1 * y) * dL_dRGB);
    SHW(1, (SH_C1 * z) * dL_dRGB);
    SHW(2, (-SH_C1 * x) * dL_dRGB);
    written = 3;
    if (deg > 1) {
      const float xx = x * x, yy = y * y, zz = z * z;
      const float xy = x * y, yz = y * z, xz = x * z;
      SHW(3, (SH_C2_0 * xy) * dL_dRGB);
      SHW(4, (SH_C2_1 * yz) * dL_dRGB);
      SHW(5, (SH_C2_2 * (2.f * zz - xx - yy)) * dL_dRGB);
      SHW(6, (SH_C2_3 * xz) * dL_dRGB);
      SHW(7, (SH_C2_4 * (xx - yy)) * dL_dRGB);
      written = 8;
      dRGBdx += SH_C2_0 * y * s3 + SH_C2_2 * 2.f * -x * s5 + SH_C2_3 * z * s6 + SH_C2_4 * 2.f * x * s7;
      dRGBdy += SH_C2_0 * x * s3 + SH_C2_1 * z * s4 + SH_C2_2 * 2.f * -y * s5 + SH_C2_4 * 2.f * -y * s7;
      dRGBdz += SH_C2_1 * y * s4 + SH_C2_2 * 2.f * 2.f * z * s5 + SH_C2_3 * x * s6;
      if (deg > 2) {
        SHW(8, (SH_C3_0 * y * (3.f * xx - yy)) * dL_dRGB);
        SHW(9, (SH_C3_1 * xy * z) * dL_dRGB);
        SHW(10, (SH_C3_2 * y * (4.f * zz - xx - yy)) * dL_dRGB);
        SHW(11, (SH_C3_3 * z * (2.f * zz - 3.f * xx - 3.f * yy)) * dL_dRGB);
        SHW(12, (SH_C3_4 * x * (4.f * zz - xx - yy)) * dL_dRGB);
        SHW(13, (SH_C3_5 * z * (xx - yy)) * dL_dRGB);
        SHW(14, (SH_C3_6 * x * (xx - 3.f * yy)) * dL_dRGB);
        written = 15;
        dRGBdx += (SH_C3_0 * s8 * 3.f * 2.f * xy + SH_C3_1 * s9 * yz + SH_C3_2 * s10 * -2.f * xy +
                   SH_C3_3 * s11 * -3.f * 2.f * xz + SH_C3_4 * s12 * (-3.f * xx + 4.f * zz - yy) +
                   SH_C3_5 * s13 * 2.f * xz + SH_C3_6 * s14 * 3.f * (xx - yy));
        dRGBdy += (SH_C3_0 * s8 * 3.f * (xx - yy) + SH_C3_1 * s9 * xz +
                   SH_C3_2 * s10 * (-3.f * yy + 4.f * zz - xx) + SH_C3_3 * s11 * -3.f * 2.f * yz +
                   SH_C3_4 * s12 * -2.f * xy + SH_C3_5 * s13 * -2.f * yz +
                   SH_C3_6 * s14 * -3.f * 2.f * xy);
        dRGBdz += (SH_C3_1 * s9 * xy + SH_C3_2 * s10 * 4.f * 2.f * yz +
                   SH_C3_3 * s11 * 3.f * (2.f * zz - xx - yy) + SH_C3_4 * s12 * 4.f * 2.f * xz +
                   SH_C3_5 * s13 * (xx - yy));
      }
    }
  }
  // coefficients above the active degree keep the zero of torch::zeros (rasterize_points.cu:204)
  for (int i = written; i < M; i++) SHW(i, (V3{0.f, 0.f, 0.f}));
#undef SHV
#undef SHW
  V3 dL_ddir = {dot(dRGBdx, dL_dRGB), dot(dRGBdy, dL_dRGB), dot(dRGBdz, dL_dRGB)};
  // auxiliary.h:118-128 dnormvdv
  const V3 vv = dir_orig, dv = dL_ddir;
  float sum2 = vv.x * vv.x + vv.y * vv.y + vv.z * vv.z;
  float invsum32 = 1.0f / sqrt(sum2 * sum2 * sum2);
  V3 r;
  r.x = ((+sum2 - vv.x * vv.x) * dv.x - vv.y * vv.x * dv.y - vv.z * vv.x * dv.z) * invsum32;
  r.y = (-vv.x * vv.y * dv.x + (sum2 - vv.y * vv.y) * dv.y - vv.z * vv.y * dv.z) * invsum32;
  r.z = (-vv.x * vv.z * dv.x - vv.y * vv.z * dv.y + (sum2 - vv.z * vv.z) * dv.z) * invsum32;
  return r;
}

constexpr int PREB_THREADS = 128;

__global__ void __launch_bounds__(PREB_THREADS, 5)
preprocess_bwd_kernel(int P, int D, int M, const float* __restrict__ means3D,
                      const int* __restrict__ radii, const float* __restrict__ dc,
                      const float* __restrict__ shs, const uint8_t* __restrict__ clamped,
                      const float* __restrict__ opacities, const float* __restrict__ scales,
                      const float* __restrict__ rotations, float scale_modifier,
                      const float* __restrict__ cov3D_precomp, const float* __restrict__ view,
                      const float* __restrict__ proj, float h_x, float h_y, float tan_fovx,
                      float tan_fovy, const float* __restrict__ campos,
                      const float* __restrict__ accum, float depth_threshold, bool antialiasing,
                      float* __restrict__ dL_dmeans3D, float* __restrict__ dL_dmeans2D,
                      float* __restrict__ dL_ddc, float* __restrict__ dL_dsh,
                      float* __restrict__ dL_dcolors_precomp, float* __restrict__ dL_dopacity_out,
                      float* __restrict__ dL_dscales, float* __restrict__ dL_drots,
                      float* __restrict__ dL_dcov3D_out, const DgrAdmmPenalty admm, bool raw_inputs,
                      double* __restrict__ pen_partial) {
  extern __shared__ float4 stage4[];
  float* stage = reinterpret_cast<float*>(stage4);
  const int first = blockIdx.x * PREB_THREADS;
  const int idx = first + threadIdx.x;
  const int count = min(PREB_THREADS, P - first);
  const bool sh_path = (dL_dsh != nullptr) && M > 0;
  const size_t nfl = (size_t)count * 3 * M;
  if (sh_path) {
    if (D > 0) {
      const float* src = shs + (size_t)first * 3 * M;
      if ((((uintptr_t)src) & 15) == 0) {
        const float4* src4 = reinterpret_cast<const float4*>(src);
        const size_t n4 = nfl >> 2;
        for (size_t i = threadIdx.x; i < n4; i += PREB_THREADS) stage4[i] = __ldg(src4 + i);
        for (size_t i = (n4 << 2) + threadIdx.x; i < nfl; i += PREB_THREADS) stage[i] = __ldg(src + i);
      } else {
        for (size_t i = threadIdx.x; i < nfl; i += PREB_THREADS) stage[i] = __ldg(src + i);
      }
    }
    __syncthreads();
  }

  double pen_xyz = 0.0, pen_dc = 0.0, pen_sh = 0.0, pen_scale = 0.0, pen_rot = 0.0, pen_op = 0.0;
  if (idx < P) {
    float* sh_row = stage + (size_t)threadIdx.x * 3 * M;
    float3 g_mean3D = {0.f, 0.f, 0.f};
    float2 g_mean2D = {0.f, 0.f};
    V3 g_dc = {0.f, 0.f, 0.f};
    float g_color[3] = {0.f, 0.f, 0.f};
    float g_opacity = 0.f;
    float g_scale[3] = {0.f, 0.f, 0.f};
    float4 g_rot = {0.f, 0.f, 0.f, 0.f};
    float g_cov3D[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const bool visible = radii[idx] > 0;
    if (!visible) {
      if (sh_path)
        for (int i = 0; i < 3 * M; i++) sh_row[i] = 0.f;
    } else {
      const float4* acc4 = reinterpret_cast<const float4*>(accum + (size_t)idx * ACC_STRIDE);
      const float4 A0 = acc4[0], A1 = acc4[1], A2 = acc4[2];
      const float2 dL_dmean2D = {A0.x, A0.y};
      float3 dL_dconic = {A0.z, A0.w, A1.x};
      float dL_dopac = A1.y;
      g_color[0] = A1.z; g_color[1] = A1.w; g_color[2] = A2.x;
      const float dL_dinvdepth = A2.y;

      const float3 mean = {means3D[3 * idx], means3D[3 * idx + 1], means3D[3 * idx + 2]};
      float3 scale = {0.f, 0.f, 0.f};
      float4 rot = {0.f, 0.f, 0.f, 0.f}, raw_rot = {0.f, 0.f, 0.f, 0.f};
      const float opac = raw_inputs ? act_sigmoid(opacities[idx]) : opacities[idx];
      float cov3D[6];
      M3 R, Mm;
      float3 s = {0.f, 0.f, 0.f};
      if (cov3D_precomp != nullptr) {
#pragma unroll
        for (int i = 0; i < 6; i++) cov3D[i] = cov3D_precomp[6 * idx + i];
      } else {
        scale = {scales[3 * idx], scales[3 * idx + 1], scales[3 * idx + 2]};
        rot = reinterpret_cast<const float4*>(rotations)[idx];
        if (raw_inputs) { raw_rot = rot; scale = act_exp3(scale); rot = act_normalize(rot); }
        // forward.cu:119-153 / backward.cu:330-352
        const float r = rot.x, x = rot.y, y = rot.z, z = rot.w;
        R = m3(1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y),
               2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x),
               2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y));
        s = {scale_modifier * scale.x, scale_modifier * scale.y, scale_modifier * scale.z};
        M3 S = m3(s.x, 0.f, 0.f, 0.f, s.y, 0.f, 0.f, 0.f, s.z);
        Mm = S * R;
        M3 Sigma = transpose(Mm) * Mm;
        cov3D[0] = Sigma.m[0][0]; cov3D[1] = Sigma.m[0][1]; cov3D[2] = Sigma.m[0][2];
        cov3D[3] = Sigma.m[1][1]; cov3D[4] = Sigma.m[1][2]; cov3D[5] = Sigma.m[2][2];
      }

      // ---- computeCov2DCUDA (backward.cu:165-325) ----
      float3 t = xform4x3(mean, view);
      const float depth = t.z;
      const float limx = 1.3f * tan_fovx;
      const float limy = 1.3f * tan_fovy;
      const float txtz = t.x / t.z;
      const float tytz = t.y / t.z;
      t.x = min(limx, max(-limx, txtz)) * t.z;
      t.y = min(limy, max(-limy, tytz)) * t.z;
      const float x_grad_mul = txtz < -limx || txtz > limx ? 0 : 1;
      const float y_grad_mul = tytz < -limy || tytz > limy ? 0 : 1;
      M3 J = m3(h_x / t.z, 0.0f, -(h_x * t.x) / (t.z * t.z), 0.0f, h_y / t.z,
                -(h_y * t.y) / (t.z * t.z), 0, 0, 0);
      M3 Wm = m3(view[0], view[4], view[8], view[1], view[5], view[9], view[2], view[6], view[10]);
      M3 Vrk = m3(cov3D[0], cov3D[1], cov3D[2], cov3D[1], cov3D[3], cov3D[4], cov3D[2], cov3D[4],
                  cov3D[5]);
      M3 Tm = Wm * J;
      M3 cov2D = transpose(Tm) * transpose(Vrk) * Tm;
      float c_xx = cov2D.m[0][0], c_xy = cov2D.m[0][1], c_yy = cov2D.m[1][1];
      constexpr float h_var = 0.3f;
      float d_inside_root = 0.f;
      if (antialiasing) {
        const float det_cov = c_xx * c_yy - c_xy * c_xy;
        c_xx += h_var;
        c_yy += h_var;
        const float det_cov_plus_h_cov = c_xx * c_yy - c_xy * c_xy;
        const float h_convolution_scaling = sqrt(max(0.000025f, det_cov / det_cov_plus_h_cov));
        const float dL_dopacity_v = dL_dopac;
        const float d_h_convolution_scaling = dL_dopacity_v * opac;
        dL_dopac = dL_dopacity_v * h_convolution_scaling;
        d_inside_root = (det_cov / det_cov_plus_h_cov) <= 0.000025f
                            ? 0.f
                            : d_h_convolution_scaling / (2 * h_convolution_scaling);
      } else {
        c_xx += h_var;
        c_yy += h_var;
      }
      float dL_dc_xx = 0, dL_dc_xy = 0, dL_dc_yy = 0;
      if (antialiasing) {
        const float x = c_xx, y = c_yy, z = c_xy, w = h_var;
        const float q = w * w + w * (x + y) + x * y - z * z;
        const float denom_f = d_inside_root / (q * q);
        dL_dc_xx = w * (w * y + y * y + z * z) * denom_f;
        dL_dc_yy = w * (w * x + x * x + z * z) * denom_f;
        dL_dc_xy = -2.f * w * z * (w + x + y) * denom_f;
      }
      const float denom = c_xx * c_yy - c_xy * c_xy;
      const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
      if (denom2inv != 0) {
        dL_dc_xx += denom2inv * (-c_yy * c_yy * dL_dconic.x + 2 * c_xy * c_yy * dL_dconic.y +
                                 (denom - c_xx * c_yy) * dL_dconic.z);
        dL_dc_yy += denom2inv * (-c_xx * c_xx * dL_dconic.z + 2 * c_xx * c_xy * dL_dconic.y +
                                 (denom - c_xx * c_yy) * dL_dconic.x);
        dL_dc_xy += denom2inv * 2 * (c_xy * c_yy * dL_dconic.x -
                                     (denom + 2 * c_xy * c_xy) * dL_dconic.y +
                                     c_xx * c_xy * dL_dconic.z);
        const float T00 = Tm.m[0][0], T01 = Tm.m[0][1], T02 = Tm.m[0][2];
        const float T10 = Tm.m[1][0], T11 = Tm.m[1][1], T12 = Tm.m[1][2];
        g_cov3D[0] = (T00 * T00 * dL_dc_xx + T00 * T10 * dL_dc_xy + T10 * T10 * dL_dc_yy);
        g_cov3D[3] = (T01 * T01 * dL_dc_xx + T01 * T11 * dL_dc_xy + T11 * T11 * dL_dc_yy);
        g_cov3D[5] = (T02 * T02 * dL_dc_xx + T02 * T12 * dL_dc_xy + T12 * T12 * dL_dc_yy);
        g_cov3D[1] = 2 * T00 * T01 * dL_dc_xx + (T00 * T11 + T01 * T10) * dL_dc_xy + 2 * T10 * T11 * dL_dc_yy;
        g_cov3D[2] = 2 * T00 * T02 * dL_dc_xx + (T00 * T12 + T02 * T10) * dL_dc_xy + 2 * T10 * T12 * dL_dc_yy;
        g_cov3D[4] = 2 * T02 * T01 * dL_dc_xx + (T01 * T12 + T02 * T11) * dL_dc_xy + 2 * T11 * T12 * dL_dc_yy;
      }
      float dL_dT[2][3];
#pragma unroll
      for (int k = 0; k < 3; k++) {
        const float a = Tm.m[0][0] * Vrk.m[k][0] + Tm.m[0][1] * Vrk.m[k][1] + Tm.m[0][2] * Vrk.m[k][2];
        const float b = Tm.m[1][0] * Vrk.m[k][0] + Tm.m[1][1] * Vrk.m[k][1] + Tm.m[1][2] * Vrk.m[k][2];
        dL_dT[0][k] = 2 * a * dL_dc_xx + b * dL_dc_xy;
        dL_dT[1][k] = 2 * b * dL_dc_yy + a * dL_dc_xy;
      }
      const float dL_dJ00 = Wm.m[0][0] * dL_dT[0][0] + Wm.m[0][1] * dL_dT[0][1] + Wm.m[0][2] * dL_dT[0][2];
      const float dL_dJ02 = Wm.m[2][0] * dL_dT[0][0] + Wm.m[2][1] * dL_dT[0][1] + Wm.m[2][2] * dL_dT[0][2];
      const float dL_dJ11 = Wm.m[1][0] * dL_dT[1][0] + Wm.m[1][1] * dL_dT[1][1] + Wm.m[1][2] * dL_dT[1][2];
      const float dL_dJ12 = Wm.m[2][0] * dL_dT[1][0] + Wm.m[2][1] * dL_dT[1][1] + Wm.m[2][2] * dL_dT[1][2];
      const float tz = 1.f / t.z;
      const float tz2 = tz * tz;
      const float tz3 = tz2 * tz;
      const float dL_dtx = x_grad_mul * -h_x * tz2 * dL_dJ02;
      const float dL_dty = y_grad_mul * -h_y * tz2 * dL_dJ12;
      const float dL_dtz = -h_x * tz2 * dL_dJ00 - h_y * tz2 * dL_dJ11 + (2 * h_x * t.x) * tz3 * dL_dJ02 +
                           (2 * h_y * t.y) * tz3 * dL_dJ12 - dL_dinvdepth * tz2;
      // transformVec4x3Transpose (auxiliary.h:100-108)
      g_mean3D = {view[0] * dL_dtx + view[1] * dL_dty + view[2] * dL_dtz,
                  view[4] * dL_dtx + view[5] * dL_dty + view[6] * dL_dtz,
                  view[8] * dL_dtx + view[9] * dL_dty + view[10] * dL_dtz};

      // ---- preprocessCUDA backward (backward.cu:421-450) ----
      float4 m_hom = xform4x4(mean, proj);
      float m_w = 1.0f / (m_hom.w + 0.0000001f);
      float mul1 = (proj[0] * mean.x + proj[4] * mean.y + proj[8] * mean.z + proj[12]) * m_w * m_w;
      float mul2 = (proj[1] * mean.x + proj[5] * mean.y + proj[9] * mean.z + proj[13]) * m_w * m_w;
      g_mean3D.x += (proj[0] * m_w - proj[3] * mul1) * dL_dmean2D.x + (proj[1] * m_w - proj[3] * mul2) * dL_dmean2D.y;
      g_mean3D.y += (proj[4] * m_w - proj[7] * mul1) * dL_dmean2D.x + (proj[5] * m_w - proj[7] * mul2) * dL_dmean2D.y;
      g_mean3D.z += (proj[8] * m_w - proj[11] * mul1) * dL_dmean2D.x + (proj[9] * m_w - proj[11] * mul2) * dL_dmean2D.y;

      if (sh_path) {
        V3 pos = {mean.x, mean.y, mean.z};
        V3 cp = {campos[0], campos[1], campos[2]};
        V3 dLc = {g_color[0], g_color[1], g_color[2]};
        V3 dm = sh_backward(D, M, pos, cp, clamped[idx], dLc, sh_row, g_dc);
        g_mean3D.x += dm.x; g_mean3D.y += dm.y; g_mean3D.z += dm.z;
      }

      if (cov3D_precomp == nullptr) {
        // computeCov3D backward (backward.cu:354-392)
        M3 dL_dSigma = m3(g_cov3D[0], 0.5f * g_cov3D[1], 0.5f * g_cov3D[2], 0.5f * g_cov3D[1],
                          g_cov3D[3], 0.5f * g_cov3D[4], 0.5f * g_cov3D[2], 0.5f * g_cov3D[4],
                          g_cov3D[5]);
        M3 dL_dM = Mm * dL_dSigma;
#pragma unroll
        for (int c = 0; c < 3; c++)
#pragma unroll
          for (int r = 0; r < 3; r++) dL_dM.m[c][r] *= 2.0f;
        M3 Rt = transpose(R);
        M3 dL_dMt = transpose(dL_dM);
        g_scale[0] = Rt.m[0][0] * dL_dMt.m[0][0] + Rt.m[0][1] * dL_dMt.m[0][1] + Rt.m[0][2] * dL_dMt.m[0][2];
        g_scale[1] = Rt.m[1][0] * dL_dMt.m[1][0] + Rt.m[1][1] * dL_dMt.m[1][1] + Rt.m[1][2] * dL_dMt.m[1][2];
        g_scale[2] = Rt.m[2][0] * dL_dMt.m[2][0] + Rt.m[2][1] * dL_dMt.m[2][1] + Rt.m[2][2] * dL_dMt.m[2][2];
#pragma unroll
        for (int r = 0; r < 3; r++) {
          dL_dMt.m[0][r] *= s.x;
          dL_dMt.m[1][r] *= s.y;
          dL_dMt.m[2][r] *= s.z;
        }
        const float r = rot.x, x = rot.y, y = rot.z, z = rot.w;
        g_rot.x = 2 * z * (dL_dMt.m[0][1] - dL_dMt.m[1][0]) + 2 * y * (dL_dMt.m[2][0] - dL_dMt.m[0][2]) + 2 * x * (dL_dMt.m[1][2] - dL_dMt.m[2][1]);
        g_rot.y = 2 * y * (dL_dMt.m[1][0] + dL_dMt.m[0][1]) + 2 * z * (dL_dMt.m[2][0] + dL_dMt.m[0][2]) + 2 * r * (dL_dMt.m[1][2] - dL_dMt.m[2][1]) - 4 * x * (dL_dMt.m[2][2] + dL_dMt.m[1][1]);
        g_rot.z = 2 * x * (dL_dMt.m[1][0] + dL_dMt.m[0][1]) + 2 * r * (dL_dMt.m[2][0] - dL_dMt.m[0][2]) + 2 * z * (dL_dMt.m[1][2] + dL_dMt.m[2][1]) - 4 * y * (dL_dMt.m[2][2] + dL_dMt.m[0][0]);
        g_rot.w = 2 * r * (dL_dMt.m[0][1] - dL_dMt.m[1][0]) + 2 * x * (dL_dMt.m[2][0] + dL_dMt.m[0][2]) + 2 * y * (dL_dMt.m[1][2] + dL_dMt.m[2][1]) - 4 * z * (dL_dMt.m[1][1] + dL_dMt.m[0][0]);
        // scale gradient is w.r.t. the given scale: dL/dscale = mod * dL/d(mod*scale)?  No — the
        // reference returns dot(Rt[i], dL_dMt[i]) unscaled by mod (backward.cu:374-377); keep it.
      }
      g_opacity = dL_dopac;
      if (raw_inputs) {
        // chain rule through the activations, in the form of torch's own backward formulas
        // (sigmoid_backward: g*(1-y)*y; exp: g*y; normalize = q / max(|q|, eps))
        g_opacity = dL_dopac * (1.0f - opac) * opac;
        if (cov3D_precomp == nullptr) {
          g_scale[0] *= scale.x; g_scale[1] *= scale.y; g_scale[2] *= scale.z;
          const float n = quat_norm(raw_rot);
          const float dot = g_rot.x * rot.x + g_rot.y * rot.y + g_rot.z * rot.z + g_rot.w * rot.w;
          const float inv_n = 1.0f / n;
          // |q| below the clamp: the denominator is the constant eps, no projection term
          const float pd = (n > 1e-12f) ? dot : 0.f;
          g_rot = {(g_rot.x - rot.x * pd) * inv_n, (g_rot.y - rot.y * pd) * inv_n,
                   (g_rot.z - rot.z * pd) * inv_n, (g_rot.w - rot.w * pd) * inv_n};
        }
      }
      g_mean2D = dL_dmean2D;
      if (depth_threshold > 0.f) {  // __init__.py:171-185
        const float q = depth / depth_threshold;
        const float sf = fminf(1.0f, q * q);
        g_mean2D.x *= sf;
        g_mean2D.y *= sf;
      }
    }
    if (admm.zmu_xyz) {
      // Fused ADMM penalty (slave_gaussian_trainer.py:161-202) on the consensus variables that are
      // direct rasterizer inputs: grad += rho_g/numel_g * (x - (z - u)), for EVERY Gaussian of the block.
      const float r0 = means3D[3 * idx] - admm.zmu_xyz[3 * idx];
      const float r1 = means3D[3 * idx + 1] - admm.zmu_xyz[3 * idx + 1];
      const float r2 = means3D[3 * idx + 2] - admm.zmu_xyz[3 * idx + 2];
      g_mean3D.x += admm.coef_xyz * r0; g_mean3D.y += admm.coef_xyz * r1; g_mean3D.z += admm.coef_xyz * r2;
      pen_xyz = (double)(r0 * r0 + r1 * r1 + r2 * r2);
      if (dL_ddc && admm.zmu_dc) {
        const float q0 = dc[3 * idx] - admm.zmu_dc[3 * idx];
        const float q1 = dc[3 * idx + 1] - admm.zmu_dc[3 * idx + 1];
        const float q2 = dc[3 * idx + 2] - admm.zmu_dc[3 * idx + 2];
        g_dc.x += admm.coef_dc * q0; g_dc.y += admm.coef_dc * q1; g_dc.z += admm.coef_dc * q2;
        pen_dc = (double)(q0 * q0 + q1 * q1 + q2 * q2);
      }
      // the remaining groups: the tensors as passed are the consensus variables (raw_inputs)
      if (dL_dscales && admm.zmu_scale) {
        const float q0 = scales[3 * idx] - admm.zmu_scale[3 * idx];
        const float q1 = scales[3 * idx + 1] - admm.zmu_scale[3 * idx + 1];
        const float q2 = scales[3 * idx + 2] - admm.zmu_scale[3 * idx + 2];
        g_scale[0] += admm.coef_scale * q0; g_scale[1] += admm.coef_scale * q1; g_scale[2] += admm.coef_scale * q2;
        pen_scale = (double)(q0 * q0 + q1 * q1 + q2 * q2);
      }
      if (dL_drots && admm.zmu_rot) {
        const float4 xr = reinterpret_cast<const float4*>(rotations)[idx];
        const float4 zr = reinterpret_cast<const float4*>(admm.zmu_rot)[idx];
        const float q0 = xr.x - zr.x, q1 = xr.y - zr.y, q2 = xr.z - zr.z, q3 = xr.w - zr.w;
        g_rot.x += admm.coef_rot * q0; g_rot.y += admm.coef_rot * q1; g_rot.z += admm.coef_rot * q2; g_rot.w += admm.coef_rot * q3;
        pen_rot = (double)(q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3);
      }
      if (admm.zmu_opacity) {
        const float q0 = opacities[idx] - admm.zmu_opacity[idx];
        g_opacity += admm.coef_opacity * q0;
        pen_op = (double)(q0 * q0);
      }
    }
    dL_dmeans3D[3 * idx] = g_mean3D.x; dL_dmeans3D[3 * idx + 1] = g_mean3D.y; dL_dmeans3D[3 * idx + 2] = g_mean3D.z;
    dL_dmeans2D[3 * idx] = g_mean2D.x; dL_dmeans2D[3 * idx + 1] = g_mean2D.y; dL_dmeans2D[3 * idx + 2] = 0.f;
    dL_dopacity_out[idx] = g_opacity;
    if (dL_ddc) { dL_ddc[3 * idx] = g_dc.x; dL_ddc[3 * idx + 1] = g_dc.y; dL_ddc[3 * idx + 2] = g_dc.z; }
    if (dL_dcolors_precomp) {
      dL_dcolors_precomp[3 * idx] = g_color[0]; dL_dcolors_precomp[3 * idx + 1] = g_color[1]; dL_dcolors_precomp[3 * idx + 2] = g_color[2];
    }
    if (dL_dscales) { dL_dscales[3 * idx] = g_scale[0]; dL_dscales[3 * idx + 1] = g_scale[1]; dL_dscales[3 * idx + 2] = g_scale[2]; }
    if (dL_drots) reinterpret_cast<float4*>(dL_drots)[idx] = g_rot;
    if (dL_dcov3D_out) {
#pragma unroll
      for (int i = 0; i < 6; i++) dL_dcov3D_out[6 * idx + i] = g_cov3D[i];
    }
  }

  if (sh_path) {
    __syncthreads();
    float* dst = dL_dsh + (size_t)first * 3 * M;
    const bool pen = admm.zmu_sh != nullptr;
    const float* xs = shs + (size_t)first * 3 * M;            // still L2-resident: this CTA just read it
    const float* zs = pen ? admm.zmu_sh + (size_t)first * 3 * M : nullptr;
    const bool al16 = ((((uintptr_t)dst) | ((uintptr_t)xs) | ((uintptr_t)zs)) & 15) == 0;
    if (al16) {
      float4* dst4 = reinterpret_cast<float4*>(dst);
      const size_t n4 = nfl >> 2;
      for (size_t i = threadIdx.x; i < n4; i += PREB_THREADS) {
        float4 g = stage4[i];
        if (pen) {
          const float4 xv = __ldg(reinterpret_cast<const float4*>(xs) + i);
          const float4 zv = __ldg(reinterpret_cast<const float4*>(zs) + i);
          const float r0 = xv.x - zv.x, r1 = xv.y - zv.y, r2 = xv.z - zv.z, r3 = xv.w - zv.w;
          g.x += admm.coef_sh * r0; g.y += admm.coef_sh * r1; g.z += admm.coef_sh * r2; g.w += admm.coef_sh * r3;
          pen_sh += (double)(r0 * r0 + r1 * r1 + r2 * r2 + r3 * r3);
        }
        dst4[i] = g;
      }
      for (size_t i = (n4 << 2) + threadIdx.x; i < nfl; i += PREB_THREADS) {
        float g = stage[i];
        if (pen) { const float r = xs[i] - zs[i]; g += admm.coef_sh * r; pen_sh += (double)(r * r); }
        dst[i] = g;
      }
    } else {
      for (size_t i = threadIdx.x; i < nfl; i += PREB_THREADS) {
        float g = stage[i];
        if (pen) { const float r = xs[i] - zs[i]; g += admm.coef_sh * r; pen_sh += (double)(r * r); }
        dst[i] = g;
      }
    }
  }
  if (pen_partial) {
    // Penalty sums for logging: one row of 6 partial sums per CTA, folded by penalty_sum_kernel.  (One
    // FP64 atomic per warp and group on six shared addresses serialised in L2: +0.2 ms per backward.)
    __shared__ double red[6][PREB_THREADS / 32];
    double v[6] = {pen_xyz, pen_dc, pen_sh, pen_scale, pen_rot, pen_op};
#pragma unroll
    for (int q = 0; q < 6; q++) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v[q] += __shfl_xor_sync(0xffffffffu, v[q], o);
      if ((threadIdx.x & 31) == 0) red[q][threadIdx.x >> 5] = v[q];
    }
    __syncthreads();
    if (threadIdx.x < 6) {
      double t = 0.0;
#pragma unroll
      for (int w = 0; w < PREB_THREADS / 32; w++) t += red[threadIdx.x][w];
      pen_partial[(size_t)blockIdx.x * 6 + threadIdx.x] = t;
    }
  }
}

// sums[g] += sum over CTAs of pen_partial[cta][g]; one CTA per group, fixed order => deterministic.
__global__ void __launch_bounds__(256)
penalty_sum_kernel(const double* __restrict__ pen_partial, int nblk, double* __restrict__ sums) {
  __shared__ double red[8];
  const int g = blockIdx.x;
  double t = 0.0;
  for (int i = threadIdx.x; i < nblk; i += 256) t += pen_partial[(size_t)i * 6 + g];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = t;
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0;
#pragma unroll
    for (int w = 0; w < 8; w++) a += red[w];
    sums[g] += a;
  }
}

size_t bwd_penalty_partial_bytes(size_t P) {
  return align_up(((P + PREB_THREADS - 1) / PREB_THREADS) * 6 * sizeof(double));
}


int launch_preprocess_bwd(const DgrParams& prm, const float* means3D, const float* dc,
                          const float* sh, const float* opacities, const float* scales,
                          const float* rotations, const float* cov3D_precomp, const int32_t* radii,
                          const GeomState& g, const float* accum, float* dL_dmeans3D,
                          float* dL_dmeans2D, float* dL_ddc, float* dL_dsh,
                          float* dL_dcolors_precomp, float* dL_dopacity, float* dL_dscales,
                          float* dL_drotations, float* dL_dcov3D, double* pen_partial,
                          cudaStream_t s) {
  const float focal_y = prm.H / (2.0f * prm.tanfovy);
  const float focal_x = prm.W / (2.0f * prm.tanfovx);
  const bool sh_path = dL_dsh != nullptr && prm.M > 0;
  DgrAdmmPenalty admm;
  memset(&admm, 0, sizeof(admm));
  if (prm.admm) admm = *prm.admm;
  if (!sh_path) admm.zmu_sh = nullptr;
  size_t smem = sh_path ? (size_t)PREB_THREADS * 3 * prm.M * sizeof(float) + 16 : 0;
  static bool attr_set[64] = {};
  allow_big_smem(preprocess_bwd_kernel, attr_set);
  if (smem > 96 * 1024) {
    set_error("preprocess_bwd: SH row too large (M=%d)", prm.M);
    return 1;
  }
  const int nblk = (prm.P + PREB_THREADS - 1) / PREB_THREADS;
  const bool sums_on = admm.zmu_xyz != nullptr && admm.sums != nullptr && pen_partial != nullptr;
  preprocess_bwd_kernel<<<nblk, PREB_THREADS, smem, s>>>(
      prm.P, prm.D, prm.M, means3D, radii, dc, sh, g.clamped, opacities, scales, rotations,
      prm.scale_modifier, cov3D_precomp, prm.viewmatrix, prm.projmatrix, focal_x, focal_y,
      prm.tanfovx, prm.tanfovy, prm.campos, accum, prm.depth_threshold, prm.antialiasing != 0,
      dL_dmeans3D, dL_dmeans2D, dL_ddc, dL_dsh, dL_dcolors_precomp, dL_dopacity, dL_dscales,
      dL_drotations, dL_dcov3D, admm, prm.raw_inputs != 0, sums_on ? pen_partial : nullptr);
  if (sums_on) {
    penalty_sum_kernel<<<6, 256, 0, s>>>(pen_partial, nblk, admm.sums);
    return check_launch("preprocess_bwd", prm.debug != 0, s, 2);
  }
  return check_launch("preprocess_bwd", prm.debug != 0, s);
}

}  // namespace dogs
