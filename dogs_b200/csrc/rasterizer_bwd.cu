// Per-Gaussian backward of the rasterizer (sm_100a): preprocess_bwd_kernel fuses computeCov2DCUDA
// (cuda_rasterizer/backward.cu:149-326) and preprocessCUDA (backward.cu:398-451) and writes every output
// gradient exactly once (zeros for culled Gaussians), which removes the 308 B/Gaussian of torch::zeros
// memsets (rasterize_points.cu:196-207).  It consumes the 12-float per-Gaussian accumulators written by
// blend_bwd_kernel (blend_bwd.cu).
//
// Precision: the chain conic -> cov2D -> cov3D -> (scale, rotation) is evaluated in FP64.  For needle-shaped
// Gaussians (axis ratios of 100 and more) it is ill-conditioned: the float32 evaluation — the reference's and
// the first version of this kernel alike — is off by up to 1e-3 of the tensor maximum at 3 M Gaussians
// against an FP64 evaluation of the same formulas (tools/gpu_truth_attrib.py: float32 accumulators + FP64
// chain rule 1e-5, float32 chain rule 6e-4).  B200 issues FP64 at half the FP32 rate and the kernel is
// HBM-bound, so the ~400 double operations per visible Gaussian hide behind the 520 B it moves.
#include <string.h>
#include "common.cuh"

namespace dogs {

constexpr int ACC_STRIDE = 12;  // floats per Gaussian in the global accumulator (3 x float4)
// accumulator slots: 0 mean2D.x, 1 mean2D.y, 2 conic.x, 3 conic.y, 4 conic.w, 5 opacity,
//                    6..8 colour rgb, 9 inverse depth, 10..11 unused

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

// ---- covariance chain: computeCov2DCUDA (backward.cu:165-325) + computeCov3D backward (:354-392) -------------
// R3 = working precision.  In: the Gaussian (mean, activated scale * modifier, normalised quaternion or a
// precomputed cov3D), the camera, and dL/dconic, dL/dopacity, dL/dinvdepth from the accumulators.
// Out: the covariance-path part of dL/dmean3D, dL/dcov3D, dL/dscale, dL/drot, the (antialiasing-corrected)
// dL/dopacity and the view-space depth.
template <typename R3> struct Mat3 { R3 m[3][3]; };   // column-major m[c][r], GLM's product order
template <typename R3>
__device__ __forceinline__ Mat3<R3> mat3(R3 a, R3 b, R3 c, R3 d, R3 e, R3 f, R3 g, R3 h, R3 i) {
  Mat3<R3> r;
  r.m[0][0] = a; r.m[0][1] = b; r.m[0][2] = c;
  r.m[1][0] = d; r.m[1][1] = e; r.m[1][2] = f;
  r.m[2][0] = g; r.m[2][1] = h; r.m[2][2] = i;
  return r;
}
template <typename R3> __device__ __forceinline__ Mat3<R3> mmul(const Mat3<R3>& A, const Mat3<R3>& B) {
  Mat3<R3> C;
#pragma unroll
  for (int c = 0; c < 3; c++)
#pragma unroll
    for (int r = 0; r < 3; r++) C.m[c][r] = A.m[0][r] * B.m[c][0] + A.m[1][r] * B.m[c][1] + A.m[2][r] * B.m[c][2];
  return C;
}
template <typename R3> __device__ __forceinline__ Mat3<R3> mtr(const Mat3<R3>& A) {
  Mat3<R3> C;
#pragma unroll
  for (int c = 0; c < 3; c++)
#pragma unroll
    for (int r = 0; r < 3; r++) C.m[c][r] = A.m[r][c];
  return C;
}

template <typename R3>
__device__ __forceinline__ void cov_chain_bwd(const float3 mean_f, const float3 scale_f, const float4 rot_f,
                                              float scale_modifier, const float* __restrict__ cov3D_pre,
                                              const float* __restrict__ view, float h_x_f, float h_y_f,
                                              float tan_fovx, float tan_fovy, bool antialiasing, float opac,
                                              const float3 dL_dconic_f, float dL_dinvdepth_f, float& dL_dopac,
                                              float3& g_mean3D, float g_cov3D[6], float g_scale[3], float4& g_rot,
                                              float& depth_out) {
  const R3 h_x = h_x_f, h_y = h_y_f;
  const R3 half = R3(0.5), two = R3(2), one = R3(1);
  R3 cov3D[6];
  const R3 sx = R3(scale_modifier) * R3(scale_f.x), sy = R3(scale_modifier) * R3(scale_f.y), sz = R3(scale_modifier) * R3(scale_f.z);
  // forward.cu:119-153 / backward.cu:330-352.  R and M = S * R are rebuilt where the cov3D backward needs them
  // again instead of being kept alive across the cov2D part (18 doubles = 36 registers).
  auto rotation = [&]() {
    const R3 r = rot_f.x, x = rot_f.y, y = rot_f.z, z = rot_f.w;
    return mat3<R3>(one - two * (y * y + z * z), two * (x * y - r * z), two * (x * z + r * y),
                    two * (x * y + r * z), one - two * (x * x + z * z), two * (y * z - r * x),
                    two * (x * z - r * y), two * (y * z + r * x), one - two * (x * x + y * y));
  };
  if (cov3D_pre != nullptr) {
#pragma unroll
    for (int i = 0; i < 6; i++) cov3D[i] = cov3D_pre[i];
  } else {
    const Mat3<R3> Mm = mmul(mat3<R3>(sx, 0, 0, 0, sy, 0, 0, 0, sz), rotation());
    const Mat3<R3> Sigma = mmul(mtr(Mm), Mm);
    cov3D[0] = Sigma.m[0][0]; cov3D[1] = Sigma.m[0][1]; cov3D[2] = Sigma.m[0][2];
    cov3D[3] = Sigma.m[1][1]; cov3D[4] = Sigma.m[1][2]; cov3D[5] = Sigma.m[2][2];
  }
  const R3 mx = mean_f.x, my = mean_f.y, mz = mean_f.z;
  R3 tx = R3(view[0]) * mx + R3(view[4]) * my + R3(view[8]) * mz + R3(view[12]);
  R3 ty = R3(view[1]) * mx + R3(view[5]) * my + R3(view[9]) * mz + R3(view[13]);
  const R3 tz = R3(view[2]) * mx + R3(view[6]) * my + R3(view[10]) * mz + R3(view[14]);
  depth_out = (float)tz;
  const R3 limx = R3(1.3f) * R3(tan_fovx), limy = R3(1.3f) * R3(tan_fovy);
  const R3 txtz = tx / tz, tytz = ty / tz;
  tx = min(limx, max(-limx, txtz)) * tz;
  ty = min(limy, max(-limy, tytz)) * tz;
  const R3 x_grad_mul = (txtz < -limx || txtz > limx) ? R3(0) : one;
  const R3 y_grad_mul = (tytz < -limy || tytz > limy) ? R3(0) : one;
  const Mat3<R3> J = mat3<R3>(h_x / tz, 0, -(h_x * tx) / (tz * tz), 0, h_y / tz, -(h_y * ty) / (tz * tz), 0, 0, 0);
  const Mat3<R3> Wm = mat3<R3>(view[0], view[4], view[8], view[1], view[5], view[9], view[2], view[6], view[10]);
  const Mat3<R3> Vrk = mat3<R3>(cov3D[0], cov3D[1], cov3D[2], cov3D[1], cov3D[3], cov3D[4], cov3D[2], cov3D[4], cov3D[5]);
  const Mat3<R3> Tm = mmul(Wm, J);
  const Mat3<R3> cov2D = mmul(mmul(mtr(Tm), mtr(Vrk)), Tm);
  R3 c_xx = cov2D.m[0][0], c_xy = cov2D.m[0][1], c_yy = cov2D.m[1][1];
  const R3 h_var = R3(0.3f);
  R3 d_inside_root = 0;
  if (antialiasing) {
    const R3 det_cov = c_xx * c_yy - c_xy * c_xy;
    c_xx += h_var;
    c_yy += h_var;
    const R3 det_cov_plus_h_cov = c_xx * c_yy - c_xy * c_xy;
    const R3 h_convolution_scaling = sqrt(max(R3(0.000025f), det_cov / det_cov_plus_h_cov));
    const R3 dL_dopacity_v = dL_dopac;
    const R3 d_h_convolution_scaling = dL_dopacity_v * R3(opac);
    dL_dopac = (float)(dL_dopacity_v * h_convolution_scaling);
    d_inside_root = (det_cov / det_cov_plus_h_cov) <= R3(0.000025f) ? R3(0)
                                                                   : d_h_convolution_scaling / (two * h_convolution_scaling);
  } else {
    c_xx += h_var;
    c_yy += h_var;
  }
  R3 dL_dc_xx = 0, dL_dc_xy = 0, dL_dc_yy = 0;
  if (antialiasing) {
    const R3 x = c_xx, y = c_yy, z = c_xy, w = h_var;
    const R3 q = w * w + w * (x + y) + x * y - z * z;
    const R3 denom_f = d_inside_root / (q * q);
    dL_dc_xx = w * (w * y + y * y + z * z) * denom_f;
    dL_dc_yy = w * (w * x + x * x + z * z) * denom_f;
    dL_dc_xy = -two * w * z * (w + x + y) * denom_f;
  }
  const R3 dcx = dL_dconic_f.x, dcy = dL_dconic_f.y, dcz = dL_dconic_f.z;
  const R3 denom = c_xx * c_yy - c_xy * c_xy;
  const R3 denom2inv = one / ((denom * denom) + R3(0.0000001f));
  R3 gc[6] = {0, 0, 0, 0, 0, 0};
  if (denom2inv != 0) {
    dL_dc_xx += denom2inv * (-c_yy * c_yy * dcx + two * c_xy * c_yy * dcy + (denom - c_xx * c_yy) * dcz);
    dL_dc_yy += denom2inv * (-c_xx * c_xx * dcz + two * c_xx * c_xy * dcy + (denom - c_xx * c_yy) * dcx);
    dL_dc_xy += denom2inv * two * (c_xy * c_yy * dcx - (denom + two * c_xy * c_xy) * dcy + c_xx * c_xy * dcz);
    const R3 T00 = Tm.m[0][0], T01 = Tm.m[0][1], T02 = Tm.m[0][2];
    const R3 T10 = Tm.m[1][0], T11 = Tm.m[1][1], T12 = Tm.m[1][2];
    gc[0] = (T00 * T00 * dL_dc_xx + T00 * T10 * dL_dc_xy + T10 * T10 * dL_dc_yy);
    gc[3] = (T01 * T01 * dL_dc_xx + T01 * T11 * dL_dc_xy + T11 * T11 * dL_dc_yy);
    gc[5] = (T02 * T02 * dL_dc_xx + T02 * T12 * dL_dc_xy + T12 * T12 * dL_dc_yy);
    gc[1] = two * T00 * T01 * dL_dc_xx + (T00 * T11 + T01 * T10) * dL_dc_xy + two * T10 * T11 * dL_dc_yy;
    gc[2] = two * T00 * T02 * dL_dc_xx + (T00 * T12 + T02 * T10) * dL_dc_xy + two * T10 * T12 * dL_dc_yy;
    gc[4] = two * T02 * T01 * dL_dc_xx + (T01 * T12 + T02 * T11) * dL_dc_xy + two * T11 * T12 * dL_dc_yy;
  }
#pragma unroll
  for (int i = 0; i < 6; i++) g_cov3D[i] = (float)gc[i];
  R3 dL_dT[2][3];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const R3 a = Tm.m[0][0] * Vrk.m[k][0] + Tm.m[0][1] * Vrk.m[k][1] + Tm.m[0][2] * Vrk.m[k][2];
    const R3 b = Tm.m[1][0] * Vrk.m[k][0] + Tm.m[1][1] * Vrk.m[k][1] + Tm.m[1][2] * Vrk.m[k][2];
    dL_dT[0][k] = two * a * dL_dc_xx + b * dL_dc_xy;
    dL_dT[1][k] = two * b * dL_dc_yy + a * dL_dc_xy;
  }
  const R3 dL_dJ00 = Wm.m[0][0] * dL_dT[0][0] + Wm.m[0][1] * dL_dT[0][1] + Wm.m[0][2] * dL_dT[0][2];
  const R3 dL_dJ02 = Wm.m[2][0] * dL_dT[0][0] + Wm.m[2][1] * dL_dT[0][1] + Wm.m[2][2] * dL_dT[0][2];
  const R3 dL_dJ11 = Wm.m[1][0] * dL_dT[1][0] + Wm.m[1][1] * dL_dT[1][1] + Wm.m[1][2] * dL_dT[1][2];
  const R3 dL_dJ12 = Wm.m[2][0] * dL_dT[1][0] + Wm.m[2][1] * dL_dT[1][1] + Wm.m[2][2] * dL_dT[1][2];
  const R3 itz = one / tz;
  const R3 tz2 = itz * itz;
  const R3 tz3 = tz2 * itz;
  const R3 dL_dtx = x_grad_mul * -h_x * tz2 * dL_dJ02;
  const R3 dL_dty = y_grad_mul * -h_y * tz2 * dL_dJ12;
  const R3 dL_dtz = -h_x * tz2 * dL_dJ00 - h_y * tz2 * dL_dJ11 + (two * h_x * tx) * tz3 * dL_dJ02 +
                    (two * h_y * ty) * tz3 * dL_dJ12 - R3(dL_dinvdepth_f) * tz2;
  // transformVec4x3Transpose (auxiliary.h:100-108)
  g_mean3D = {(float)(R3(view[0]) * dL_dtx + R3(view[1]) * dL_dty + R3(view[2]) * dL_dtz),
              (float)(R3(view[4]) * dL_dtx + R3(view[5]) * dL_dty + R3(view[6]) * dL_dtz),
              (float)(R3(view[8]) * dL_dtx + R3(view[9]) * dL_dty + R3(view[10]) * dL_dtz)};
  if (cov3D_pre == nullptr) {
    // computeCov3D backward (backward.cu:354-392).  The scale gradient is dot(Rt[i], dL_dMt[i]), not
    // multiplied by the modifier, as in the reference (backward.cu:374-377).
    const Mat3<R3> dL_dSigma = mat3<R3>(gc[0], half * gc[1], half * gc[2], half * gc[1], gc[3], half * gc[4],
                                        half * gc[2], half * gc[4], gc[5]);
    const Mat3<R3> Rm = rotation();
    const Mat3<R3> Mm = mmul(mat3<R3>(sx, 0, 0, 0, sy, 0, 0, 0, sz), Rm);
    Mat3<R3> dL_dM = mmul(Mm, dL_dSigma);
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
      for (int r = 0; r < 3; r++) dL_dM.m[c][r] *= two;
    const Mat3<R3> Rt = mtr(Rm);
    Mat3<R3> dMt = mtr(dL_dM);
    g_scale[0] = (float)(Rt.m[0][0] * dMt.m[0][0] + Rt.m[0][1] * dMt.m[0][1] + Rt.m[0][2] * dMt.m[0][2]);
    g_scale[1] = (float)(Rt.m[1][0] * dMt.m[1][0] + Rt.m[1][1] * dMt.m[1][1] + Rt.m[1][2] * dMt.m[1][2]);
    g_scale[2] = (float)(Rt.m[2][0] * dMt.m[2][0] + Rt.m[2][1] * dMt.m[2][1] + Rt.m[2][2] * dMt.m[2][2]);
#pragma unroll
    for (int r = 0; r < 3; r++) {
      dMt.m[0][r] *= sx;
      dMt.m[1][r] *= sy;
      dMt.m[2][r] *= sz;
    }
    const R3 r = rot_f.x, x = rot_f.y, y = rot_f.z, z = rot_f.w;
    const R3 four = R3(4);
    g_rot.x = (float)(two * z * (dMt.m[0][1] - dMt.m[1][0]) + two * y * (dMt.m[2][0] - dMt.m[0][2]) + two * x * (dMt.m[1][2] - dMt.m[2][1]));
    g_rot.y = (float)(two * y * (dMt.m[1][0] + dMt.m[0][1]) + two * z * (dMt.m[2][0] + dMt.m[0][2]) + two * r * (dMt.m[1][2] - dMt.m[2][1]) - four * x * (dMt.m[2][2] + dMt.m[1][1]));
    g_rot.z = (float)(two * x * (dMt.m[1][0] + dMt.m[0][1]) + two * r * (dMt.m[2][0] - dMt.m[0][2]) + two * z * (dMt.m[1][2] + dMt.m[2][1]) - four * y * (dMt.m[2][2] + dMt.m[0][0]));
    g_rot.w = (float)(two * r * (dMt.m[0][1] - dMt.m[1][0]) + two * x * (dMt.m[2][0] + dMt.m[0][2]) + two * y * (dMt.m[1][2] + dMt.m[2][1]) - four * z * (dMt.m[1][1] + dMt.m[0][0]));
  }
}

#ifndef DOGS_PREB_REAL
#define DOGS_PREB_REAL double
#endif
#ifndef PREB_THREADS_N
#define PREB_THREADS_N 128
#endif
#ifndef PREB_MIN_CTAS
#define PREB_MIN_CTAS 5
#endif
constexpr int PREB_THREADS = PREB_THREADS_N;

__global__ void __launch_bounds__(PREB_THREADS, PREB_MIN_CTAS)
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
  extern __shared__ __align__(128) float4 stage4[];
  float* stage = reinterpret_cast<float*>(stage4);
  __shared__ __align__(8) uint64_t sh_bar;
  const int first = blockIdx.x * PREB_THREADS;
  const int idx = first + threadIdx.x;
  const int count = min(PREB_THREADS, P - first);
  const bool sh_path = (dL_dsh != nullptr) && M > 0;
  const size_t nfl = (size_t)count * 3 * M;
  // The CTA's SH slice arrives as one bulk copy (TMA engine) while the threads run the covariance chain; the
  // gradient slice that replaces it in shared memory leaves the same way.  Fallback (slice not a multiple of
  // 16 bytes, misaligned tensor): coalesced LSU loads / stores.
  bool bulk = false;
  if (sh_path) {
    if (D > 0) {
      const float* src = shs + (size_t)first * 3 * M;
      bulk = (((uintptr_t)src) & 15) == 0 && ((nfl * sizeof(float)) & 15) == 0;
      if (bulk) {
        if (threadIdx.x == 0) {
          mbar_init(&sh_bar, 1);
          bulk_load(stage, src, (uint32_t)(nfl * sizeof(float)), &sh_bar);
        }
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
    // Every per-Gaussian input is requested here, before the first of them is needed.  (ncu, stall samples by
    // source line: the visibility test, the accumulator / mean / opacity loads, the scale / rotation loads inside
    // their `if`, and the clamp flags inside the SH branch were four memory round trips in a row — loads are not
    // moved across the branches that separate them.  The asm statements are volatile so that they stay put.)
    int rad_in;
    float4 A0, A1, A2, rot_in = {0.f, 0.f, 0.f, 0.f};
    float3 mean_in, scale_in = {0.f, 0.f, 0.f};
    float opac_in;
    uint32_t clamp_in;
    {
      const float4* acc4 = reinterpret_cast<const float4*>(accum + (size_t)idx * ACC_STRIDE);
      asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(rad_in) : "l"(radii + idx));
      asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(A0.x), "=f"(A0.y), "=f"(A0.z), "=f"(A0.w) : "l"(acc4));
      asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(A1.x), "=f"(A1.y), "=f"(A1.z), "=f"(A1.w) : "l"(acc4 + 1));
      asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(A2.x), "=f"(A2.y), "=f"(A2.z), "=f"(A2.w) : "l"(acc4 + 2));
      asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(mean_in.x) : "l"(means3D + 3 * idx));
      asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(mean_in.y) : "l"(means3D + 3 * idx + 1));
      asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(mean_in.z) : "l"(means3D + 3 * idx + 2));
      asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(opac_in) : "l"(opacities + idx));
      asm volatile("ld.global.nc.u8 %0, [%1];" : "=r"(clamp_in) : "l"(clamped + idx));
      if (cov3D_precomp == nullptr) {
        asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(scale_in.x) : "l"(scales + 3 * idx));
        asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(scale_in.y) : "l"(scales + 3 * idx + 1));
        asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(scale_in.z) : "l"(scales + 3 * idx + 2));
        asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(rot_in.x), "=f"(rot_in.y), "=f"(rot_in.z), "=f"(rot_in.w)
                     : "l"(reinterpret_cast<const float4*>(rotations) + idx));
      }
    }
    const bool visible = rad_in > 0;
    if (!visible) {
      if (sh_path) {
        if (bulk) mbar_wait(&sh_bar, 0);     // the incoming copy must not land on top of the zeros
        for (int i = 0; i < 3 * M; i++) sh_row[i] = 0.f;
      }
    } else {
      const float2 dL_dmean2D = {A0.x, A0.y};
      float3 dL_dconic = {A0.z, A0.w, A1.x};
      float dL_dopac = A1.y;
      g_color[0] = A1.z; g_color[1] = A1.w; g_color[2] = A2.x;
      const float dL_dinvdepth = A2.y;

      const float3 mean = mean_in;
      float3 scale = scale_in;
      float4 rot = rot_in, raw_rot = {0.f, 0.f, 0.f, 0.f};
      const float opac = raw_inputs ? act_sigmoid(opac_in) : opac_in;
      if (cov3D_precomp == nullptr) {
        if (raw_inputs) { raw_rot = rot; scale = act_exp3(scale); rot = act_normalize(rot); }
      }
      float depth;
      cov_chain_bwd<DOGS_PREB_REAL>(mean, scale, rot, scale_modifier, cov3D_precomp ? cov3D_precomp + 6 * idx : nullptr,
                                    view, h_x, h_y, tan_fovx, tan_fovy, antialiasing, opac, dL_dconic, dL_dinvdepth,
                                    dL_dopac, g_mean3D, g_cov3D, g_scale, g_rot, depth);

      // ---- preprocessCUDA backward (backward.cu:421-450) ----
      float4 m_hom = xform4x4(mean, proj);
      float m_w = 1.0f / (m_hom.w + 0.0000001f);
      float mul1 = (proj[0] * mean.x + proj[4] * mean.y + proj[8] * mean.z + proj[12]) * m_w * m_w;
      float mul2 = (proj[1] * mean.x + proj[5] * mean.y + proj[9] * mean.z + proj[13]) * m_w * m_w;
      g_mean3D.x += (proj[0] * m_w - proj[3] * mul1) * dL_dmean2D.x + (proj[1] * m_w - proj[3] * mul2) * dL_dmean2D.y;
      g_mean3D.y += (proj[4] * m_w - proj[7] * mul1) * dL_dmean2D.x + (proj[5] * m_w - proj[7] * mul2) * dL_dmean2D.y;
      g_mean3D.z += (proj[8] * m_w - proj[11] * mul1) * dL_dmean2D.x + (proj[9] * m_w - proj[11] * mul2) * dL_dmean2D.y;

      if (sh_path) {
        if (bulk) mbar_wait(&sh_bar, 0);
        V3 pos = {mean.x, mean.y, mean.z};
        V3 cp = {campos[0], campos[1], campos[2]};
        V3 dLc = {g_color[0], g_color[1], g_color[2]};
        V3 dm = sh_backward(D, M, pos, cp, clamp_in, dLc, sh_row, g_dc);
        g_mean3D.x += dm.x; g_mean3D.y += dm.y; g_mean3D.z += dm.z;
      } else if (dL_ddc != nullptr) {
        // degree-0 model without a rest tensor (M == 0): the colour is SH_C0 * dc + 0.5 and only dc gets a gradient
        // (the reference's `if (shs)` branch, backward.cu:445-447, is never taken for it and its dc never trains)
        const uint32_t cl = clamp_in;
        g_dc = {(cl & 1u) ? 0.f : SH_C0 * g_color[0], (cl & 2u) ? 0.f : SH_C0 * g_color[1], (cl & 4u) ? 0.f : SH_C0 * g_color[2]};
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
      // Every load of the five groups is requested before the first value is used, in ONE basic block (predicated
      // loads instead of one `if` per group): ncu showed five waits for memory in a row here, one per group —
      // 16 % of this kernel's stall samples — because loads are not moved across the groups' branches.
      const bool p_dc = dL_ddc && admm.zmu_dc, p_sc = dL_dscales && admm.zmu_scale;
      const bool p_rot = dL_drots && admm.zmu_rot, p_op = admm.zmu_opacity != nullptr;
      const float zx0 = admm.zmu_xyz[3 * idx], zx1 = admm.zmu_xyz[3 * idx + 1], zx2 = admm.zmu_xyz[3 * idx + 2];
      const float zd0 = p_dc ? admm.zmu_dc[3 * idx] : 0.f, zd1 = p_dc ? admm.zmu_dc[3 * idx + 1] : 0.f,
                  zd2 = p_dc ? admm.zmu_dc[3 * idx + 2] : 0.f;
      const float zs0 = p_sc ? admm.zmu_scale[3 * idx] : 0.f, zs1 = p_sc ? admm.zmu_scale[3 * idx + 1] : 0.f,
                  zs2 = p_sc ? admm.zmu_scale[3 * idx + 2] : 0.f;
      const float4 zero4 = {0.f, 0.f, 0.f, 0.f};
      const float4 zr = p_rot ? reinterpret_cast<const float4*>(admm.zmu_rot)[idx] : zero4;
      const float zo = p_op ? admm.zmu_opacity[idx] : 0.f;
      const float x0 = means3D[3 * idx], x1 = means3D[3 * idx + 1], x2 = means3D[3 * idx + 2];
      const float d0 = p_dc ? dc[3 * idx] : 0.f, d1 = p_dc ? dc[3 * idx + 1] : 0.f, d2 = p_dc ? dc[3 * idx + 2] : 0.f;
      const float s0 = p_sc ? scales[3 * idx] : 0.f, s1 = p_sc ? scales[3 * idx + 1] : 0.f, s2 = p_sc ? scales[3 * idx + 2] : 0.f;
      const float4 xr = p_rot ? reinterpret_cast<const float4*>(rotations)[idx] : zero4;
      const float xo = p_op ? opacities[idx] : 0.f;
      {
        const float r0 = x0 - zx0, r1 = x1 - zx1, r2 = x2 - zx2;
        g_mean3D.x += admm.coef_xyz * r0; g_mean3D.y += admm.coef_xyz * r1; g_mean3D.z += admm.coef_xyz * r2;
        pen_xyz = (double)(r0 * r0 + r1 * r1 + r2 * r2);
      }
      if (p_dc) {
        const float q0 = d0 - zd0, q1 = d1 - zd1, q2 = d2 - zd2;
        g_dc.x += admm.coef_dc * q0; g_dc.y += admm.coef_dc * q1; g_dc.z += admm.coef_dc * q2;
        pen_dc = (double)(q0 * q0 + q1 * q1 + q2 * q2);
      }
      // the remaining groups: the tensors as passed are the consensus variables (raw_inputs)
      if (p_sc) {
        const float q0 = s0 - zs0, q1 = s1 - zs1, q2 = s2 - zs2;
        g_scale[0] += admm.coef_scale * q0; g_scale[1] += admm.coef_scale * q1; g_scale[2] += admm.coef_scale * q2;
        pen_scale = (double)(q0 * q0 + q1 * q1 + q2 * q2);
      }
      if (p_rot) {
        const float q0 = xr.x - zr.x, q1 = xr.y - zr.y, q2 = xr.z - zr.z, q3 = xr.w - zr.w;
        g_rot.x += admm.coef_rot * q0; g_rot.y += admm.coef_rot * q1; g_rot.z += admm.coef_rot * q2; g_rot.w += admm.coef_rot * q3;
        pen_rot = (double)(q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3);
      }
      if (p_op) {
        const float q0 = xo - zo;
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
    if (bulk) mbar_wait(&sh_bar, 0);       // threads beyond P: nobody passes while the copy may still be in flight
    float* dst = dL_dsh + (size_t)first * 3 * M;
    const bool pen = admm.zmu_sh != nullptr;
    if (pen) {
      // ADMM penalty of the SH coefficients, added to the staged gradient slice in place: x is re-read (L2: this
      // CTA just loaded it), z - u comes from HBM.  Four 16-byte vectors per thread and array are requested before
      // the first one is used: the rolled loop of the first version waited for memory 11 times per CTA and made
      // this kernel 0.26 ms slower with the penalty than without (0.11 ms of it is the extra traffic).
      __syncthreads();                     // every thread's gradient row is in the slice
      const float* xs = shs + (size_t)first * 3 * M;
      const float* zs = admm.zmu_sh + (size_t)first * 3 * M;
      const bool al16 = ((((uintptr_t)xs) | ((uintptr_t)zs)) & 15) == 0;
      const size_t n4 = al16 ? nfl >> 2 : 0;
      const float4* x4 = reinterpret_cast<const float4*>(xs);
      const float4* z4 = reinterpret_cast<const float4*>(zs);
      float acc = 0.f;
      // (PEN_BATCH vectors per thread and array in flight: the 11.25 vectors a thread owns of a degree-3 slice take
      //  two rounds of 6 — with rounds of 4 the third wait for memory was 5 % of the kernel)
      constexpr int PEN_BATCH = 6;
      for (size_t base = 0; base < n4; base += PREB_THREADS * PEN_BATCH) {
        float4 xv[PEN_BATCH], zv[PEN_BATCH];
#pragma unroll
        for (int k = 0; k < PEN_BATCH; k++) {
          const size_t i = base + k * PREB_THREADS + threadIdx.x;
          if (i < n4) { xv[k] = __ldg(x4 + i); zv[k] = __ldg(z4 + i); }
        }
#pragma unroll
        for (int k = 0; k < PEN_BATCH; k++) {
          const size_t i = base + k * PREB_THREADS + threadIdx.x;
          if (i < n4) {
            float4 g = stage4[i];
            const float r0 = xv[k].x - zv[k].x, r1 = xv[k].y - zv[k].y, r2 = xv[k].z - zv[k].z, r3 = xv[k].w - zv[k].w;
            g.x += admm.coef_sh * r0; g.y += admm.coef_sh * r1; g.z += admm.coef_sh * r2; g.w += admm.coef_sh * r3;
            acc += r0 * r0 + r1 * r1 + r2 * r2 + r3 * r3;
            stage4[i] = g;
          }
        }
        pen_sh += (double)acc;             // <= 24 squares per step in float, everything above in double
        acc = 0.f;
      }
      for (size_t i = (n4 << 2) + threadIdx.x; i < nfl; i += PREB_THREADS) {
        const float r = xs[i] - zs[i];
        stage[i] += admm.coef_sh * r;
        pen_sh += (double)(r * r);
      }
    }
    const bool bulk_out = (((uintptr_t)dst) & 15) == 0 && ((nfl * sizeof(float)) & 15) == 0;
    if (bulk_out) fence_smem_to_async();
    __syncthreads();
    if (bulk_out) {
      if (threadIdx.x == 0) bulk_store_and_wait(dst, stage, (uint32_t)(nfl * sizeof(float)));
    } else if ((((uintptr_t)dst) & 15) == 0) {
      float4* dst4 = reinterpret_cast<float4*>(dst);
      const size_t n4 = nfl >> 2;
      for (size_t i = threadIdx.x; i < n4; i += PREB_THREADS) dst4[i] = stage4[i];
      for (size_t i = (n4 << 2) + threadIdx.x; i < nfl; i += PREB_THREADS) dst[i] = stage[i];
    } else {
      for (size_t i = threadIdx.x; i < nfl; i += PREB_THREADS) dst[i] = stage[i];
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
