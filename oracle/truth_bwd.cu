// FP64 "truth" of the rasterizer backward — TEST INFRASTRUCTURE ONLY (tests/ load libtruth.so through
// ctypes; nothing under dogs_b200/ links or calls it).
//
// Purpose: at the full C2 / C3 sizes both float32 backward implementations (the reference's, with
// float atomics in scheduling order, and this repository's) differ from each other by ~1e-4 of the
// tensor maximum on a handful of ill-conditioned Gaussians.  To decide which one is closer to the
// mathematically exact gradient of the SAME forward pass, this file evaluates the reference's
// backward formulas in double precision from the float32 forward state (which is bit-identical in
// both implementations):
//   truth_blend_bwd       cuda_rasterizer/backward.cu:453-658 (PerGaussianRenderCUDA), incl. the doubled
//                         background term; every product and every sum in FP64
//   truth_preprocess_bwd  backward.cu:149-326 (computeCov2DCUDA) + :398-451 (preprocessCUDA) + :23-144 (SH)
//                         + :330-393 (computeCov3D), inputs promoted to FP64
// The set of contributing (pixel, splat) pairs is the forward's: index below n_contrib and the forward's own
// float32 power / alpha tests.
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

struct Rec { float4 a, b, c; };   // layout of dogs::SplatRecord (csrc/common.cuh): {conic xyz, opacity} {x, y, 1/depth, r} {g, b, depth, thr}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// one CTA of 256 threads per 16x16 tile, one pixel per thread; accum[P][10] doubles:
// 0 mean2D.x, 1 mean2D.y, 2 conic.x, 3 conic.y, 4 conic.w, 5 opacity, 6..8 rgb, 9 inverse depth
__global__ void __launch_bounds__(256)
truth_blend_bwd_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list, int W, int H,
                       const float* __restrict__ bg, const Rec* __restrict__ rec,
                       const uint32_t* __restrict__ n_contrib, const float* __restrict__ dL_dpix,
                       const float* __restrict__ dL_dinvd, double* __restrict__ accum) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // warp -> 8x4 block like the forward kernel (keeps a warp's pixels compact)
  const uint32_t px = blockIdx.x * 16 + (warp & 1) * 8 + (lane & 7);
  const uint32_t py = blockIdx.y * 16 + (warp >> 1) * 4 + (lane >> 3);
  const bool inside = px < (uint32_t)W && py < (uint32_t)H;
  const uint32_t pix = py * W + px;
  const size_t HW = (size_t)W * H;
  const uint2 range = ranges[blockIdx.y * ((W + 15) / 16) + blockIdx.x];
  const uint32_t last = inside ? n_contrib[pix] : 0u;
  const float pxf = (float)px, pyf = (float)py;
  double dC[3] = {0, 0, 0}, dD = 0;
  if (inside) {
    dC[0] = dL_dpix[pix]; dC[1] = dL_dpix[HW + pix]; dC[2] = dL_dpix[2 * HW + pix];
    dD = dL_dinvd ? (double)dL_dinvd[pix] : 0.0;
  }
  uint32_t wlast = last;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) wlast = max(wlast, __shfl_xor_sync(0xffffffffu, wlast, o));

  // pass 1: totals in FP64 over the pairs the forward blended
  double Ct[3] = {0, 0, 0}, Dt = 0, T = 1.0;
  for (uint32_t e = 0; e < wlast; e++) {
    const Rec r = rec[point_list[range.x + e]];
    const float dxf = r.b.x - pxf, dyf = r.b.y - pyf;
    const float power = -0.5f * (r.a.x * dxf * dxf + r.a.z * dyf * dyf) - r.a.y * dxf * dyf;   // forward.cu:448-452
    bool c = e < last && !(power > 0.0f);
    if (c) c = !(fminf(0.99f, r.a.w * expf(power)) < 1.0f / 255.0f);
    if (!c) continue;
    const double dx = (double)r.b.x - pxf, dy = (double)r.b.y - pyf;
    const double pw = -0.5 * ((double)r.a.x * dx * dx + (double)r.a.z * dy * dy) - (double)r.a.y * dx * dy;
    const double alpha = fmin(0.99, (double)r.a.w * exp(pw));
    const double w = alpha * T;
    Ct[0] += w * r.b.w; Ct[1] += w * r.c.x; Ct[2] += w * r.c.y; Dt += w * r.b.z;
    T *= 1.0 - alpha;
  }
  const double Tfinal = T;
  const double bgdot = (double)bg[0] * dC[0] + (double)bg[1] * dC[1] + (double)bg[2] * dC[2];
  const double pixC[3] = {Ct[0] + Tfinal * bg[0], Ct[1] + Tfinal * bg[1], Ct[2] + Tfinal * bg[2]};

  // pass 2
  double Cp[3] = {0, 0, 0}, Dp = 0;
  T = 1.0;
  for (uint32_t e = 0; e < wlast; e++) {
    const uint32_t id = point_list[range.x + e];
    const Rec r = rec[id];
    const float dxf = r.b.x - pxf, dyf = r.b.y - pyf;
    const float power = -0.5f * (r.a.x * dxf * dxf + r.a.z * dyf * dyf) - r.a.y * dxf * dyf;
    bool c = e < last && !(power > 0.0f);
    if (c) c = !(fminf(0.99f, r.a.w * expf(power)) < 1.0f / 255.0f);
    if (!__any_sync(0xffffffffu, c)) continue;
    double v[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (c) {
      const double dx = (double)r.b.x - pxf, dy = (double)r.b.y - pyf;
      const double cx = r.a.x, cy = r.a.y, cz = r.a.z, o = r.a.w;
      const double pw = -0.5 * (cx * dx * dx + cz * dy * dy) - cy * dx * dy;
      const double G = exp(pw);
      const double alpha = fmin(0.99, o * G);
      const double w = alpha * T;
      const double col[3] = {r.b.w, r.c.x, r.c.y};
      const double invd = r.b.z;
      const double om = 1.0 - alpha;
      // colour / depth behind this splat, as the reference carries it (-ar, -ard): includes Tfinal*bg
      double dL_dalpha = 0.0;
#pragma unroll
      for (int ch = 0; ch < 3; ch++) {
        const double behind = pixC[ch] - Cp[ch] - w * col[ch];
        dL_dalpha += (col[ch] * T - behind / om) * dC[ch];
        Cp[ch] += w * col[ch];
      }
      {
        const double behind = Dt - Dp - w * invd;
        dL_dalpha += (invd * T - behind / om) * dD;
        Dp += w * invd;
      }
      dL_dalpha += (-Tfinal / om) * bgdot;                       // backward.cu:604-621 (bg counted twice)
      T *= om;
      const double dL_dG = o * dL_dalpha;
      const double gdx = G * dx, gdy = G * dy;
      const double dG_ddelx = -gdx * cx - gdy * cy;
      const double dG_ddely = -gdy * cz - gdx * cy;
      v[0] = dL_dG * dG_ddelx * (0.5 * W);
      v[1] = dL_dG * dG_ddely * (0.5 * H);
      v[2] = -0.5 * gdx * dx * dL_dG;
      v[3] = -0.5 * gdx * dy * dL_dG;
      v[4] = -0.5 * gdy * dy * dL_dG;
      v[5] = G * dL_dalpha;
      v[6] = w * dC[0]; v[7] = w * dC[1]; v[8] = w * dC[2];
      v[9] = w * dD;
    }
#pragma unroll
    for (int q = 0; q < 10; q++) {
      const double t = warp_sum(v[q]);
      if (lane == 0 && t != 0.0) atomicAdd(&accum[(size_t)id * 10 + q], t);
    }
  }
}

struct D3 { double x, y, z; };
struct DM3 { double m[3][3]; };   // column-major m[c][r] like GLM
__device__ __forceinline__ DM3 dm3(double a, double b, double c, double d, double e, double f, double g, double h, double i) {
  DM3 r;
  r.m[0][0] = a; r.m[0][1] = b; r.m[0][2] = c; r.m[1][0] = d; r.m[1][1] = e; r.m[1][2] = f; r.m[2][0] = g; r.m[2][1] = h; r.m[2][2] = i;
  return r;
}
__device__ __forceinline__ DM3 mul(const DM3& A, const DM3& B) {
  DM3 R;
  for (int c = 0; c < 3; c++)
    for (int r = 0; r < 3; r++) R.m[c][r] = A.m[0][r] * B.m[c][0] + A.m[1][r] * B.m[c][1] + A.m[2][r] * B.m[c][2];
  return R;
}
__device__ __forceinline__ DM3 tr(const DM3& A) {
  DM3 R;
  for (int c = 0; c < 3; c++)
    for (int r = 0; r < 3; r++) R.m[c][r] = A.m[r][c];
  return R;
}

#define C0 0.28209479177387814
#define C1 0.4886025119029199
__constant__ double C2[5] = {1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396};
__constant__ double C3[7] = {-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154,
                             -0.4570457994644658, 1.445305721320277, -0.5900435899266435};

// One thread per Gaussian.  Outputs (double): g_mean3D[P,3], g_mean2D[P,2], g_dc[P,3], g_sh[P,M,3], g_opacity[P],
// g_scale[P,3], g_rot[P,4]; zero for radii <= 0.
__global__ void __launch_bounds__(128)
truth_preprocess_bwd_kernel(int P, int D, int M, const float* __restrict__ means3D, const int* __restrict__ radii,
                            const float* __restrict__ shs, const uint8_t* __restrict__ clamped,
                            const float* __restrict__ scales, const float* __restrict__ rots, double scale_modifier,
                            const float* __restrict__ view, const float* __restrict__ proj, double h_x, double h_y,
                            double tan_fovx, double tan_fovy, const float* __restrict__ campos,
                            const double* __restrict__ accum, double depth_threshold,
                            double* __restrict__ g_mean3D, double* __restrict__ g_mean2D, double* __restrict__ g_dc,
                            double* __restrict__ g_sh, double* __restrict__ g_opacity, double* __restrict__ g_scale,
                            double* __restrict__ g_rot) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= P) return;
  for (int k = 0; k < 3; k++) { g_mean3D[3 * idx + k] = 0; g_dc[3 * idx + k] = 0; g_scale[3 * idx + k] = 0; }
  g_mean2D[2 * idx] = g_mean2D[2 * idx + 1] = 0;
  g_opacity[idx] = 0;
  for (int k = 0; k < 4; k++) g_rot[4 * idx + k] = 0;
  for (int k = 0; k < 3 * M; k++) g_sh[(size_t)idx * 3 * M + k] = 0;
  if (radii[idx] <= 0) return;
  const double* A = accum + (size_t)idx * 10;
  const double dL_dmean2D[2] = {A[0], A[1]};
  const double dL_dconic[3] = {A[2], A[3], A[4]};      // x, y, w
  const double dL_dcol[3] = {A[6], A[7], A[8]};
  const double dL_dinvdepth = A[9];
  g_opacity[idx] = A[5];

  const D3 mean = {means3D[3 * idx], means3D[3 * idx + 1], means3D[3 * idx + 2]};
  const double sx = scale_modifier * scales[3 * idx], sy = scale_modifier * scales[3 * idx + 1], sz = scale_modifier * scales[3 * idx + 2];
  const double r = rots[4 * idx], x = rots[4 * idx + 1], y = rots[4 * idx + 2], z = rots[4 * idx + 3];
  const DM3 R = dm3(1.0 - 2.0 * (y * y + z * z), 2.0 * (x * y - r * z), 2.0 * (x * z + r * y),
                    2.0 * (x * y + r * z), 1.0 - 2.0 * (x * x + z * z), 2.0 * (y * z - r * x),
                    2.0 * (x * z - r * y), 2.0 * (y * z + r * x), 1.0 - 2.0 * (x * x + y * y));
  const DM3 S = dm3(sx, 0, 0, 0, sy, 0, 0, 0, sz);
  const DM3 Mm = mul(S, R);
  const DM3 Sigma = mul(tr(Mm), Mm);
  const double cov3D[6] = {Sigma.m[0][0], Sigma.m[0][1], Sigma.m[0][2], Sigma.m[1][1], Sigma.m[1][2], Sigma.m[2][2]};

  // computeCov2DCUDA (backward.cu:165-325)
  double tx = view[0] * mean.x + view[4] * mean.y + view[8] * mean.z + view[12];
  double ty = view[1] * mean.x + view[5] * mean.y + view[9] * mean.z + view[13];
  const double tz = view[2] * mean.x + view[6] * mean.y + view[10] * mean.z + view[14];
  const double depth = tz;
  const double limx = 1.3 * tan_fovx, limy = 1.3 * tan_fovy;
  const double txtz = tx / tz, tytz = ty / tz;
  tx = fmin(limx, fmax(-limx, txtz)) * tz;
  ty = fmin(limy, fmax(-limy, tytz)) * tz;
  const double x_grad_mul = (txtz < -limx || txtz > limx) ? 0 : 1;
  const double y_grad_mul = (tytz < -limy || tytz > limy) ? 0 : 1;
  const DM3 J = dm3(h_x / tz, 0, -(h_x * tx) / (tz * tz), 0, h_y / tz, -(h_y * ty) / (tz * tz), 0, 0, 0);
  const DM3 Wm = dm3(view[0], view[4], view[8], view[1], view[5], view[9], view[2], view[6], view[10]);
  const DM3 Vrk = dm3(cov3D[0], cov3D[1], cov3D[2], cov3D[1], cov3D[3], cov3D[4], cov3D[2], cov3D[4], cov3D[5]);
  const DM3 Tm = mul(Wm, J);
  const DM3 cov2D = mul(mul(tr(Tm), tr(Vrk)), Tm);
  const double c_xx = cov2D.m[0][0] + 0.3, c_xy = cov2D.m[0][1], c_yy = cov2D.m[1][1] + 0.3;
  double dL_dc_xx = 0, dL_dc_xy = 0, dL_dc_yy = 0;
  double g_cov[6] = {0, 0, 0, 0, 0, 0};
  const double denom = c_xx * c_yy - c_xy * c_xy;
  const double denom2inv = 1.0 / ((denom * denom) + 0.0000001);
  if (denom2inv != 0) {
    dL_dc_xx = denom2inv * (-c_yy * c_yy * dL_dconic[0] + 2 * c_xy * c_yy * dL_dconic[1] + (denom - c_xx * c_yy) * dL_dconic[2]);
    dL_dc_yy = denom2inv * (-c_xx * c_xx * dL_dconic[2] + 2 * c_xx * c_xy * dL_dconic[1] + (denom - c_xx * c_yy) * dL_dconic[0]);
    dL_dc_xy = denom2inv * 2 * (c_xy * c_yy * dL_dconic[0] - (denom + 2 * c_xy * c_xy) * dL_dconic[1] + c_xx * c_xy * dL_dconic[2]);
    const double T00 = Tm.m[0][0], T01 = Tm.m[0][1], T02 = Tm.m[0][2], T10 = Tm.m[1][0], T11 = Tm.m[1][1], T12 = Tm.m[1][2];
    g_cov[0] = T00 * T00 * dL_dc_xx + T00 * T10 * dL_dc_xy + T10 * T10 * dL_dc_yy;
    g_cov[3] = T01 * T01 * dL_dc_xx + T01 * T11 * dL_dc_xy + T11 * T11 * dL_dc_yy;
    g_cov[5] = T02 * T02 * dL_dc_xx + T02 * T12 * dL_dc_xy + T12 * T12 * dL_dc_yy;
    g_cov[1] = 2 * T00 * T01 * dL_dc_xx + (T00 * T11 + T01 * T10) * dL_dc_xy + 2 * T10 * T11 * dL_dc_yy;
    g_cov[2] = 2 * T00 * T02 * dL_dc_xx + (T00 * T12 + T02 * T10) * dL_dc_xy + 2 * T10 * T12 * dL_dc_yy;
    g_cov[4] = 2 * T02 * T01 * dL_dc_xx + (T01 * T12 + T02 * T11) * dL_dc_xy + 2 * T11 * T12 * dL_dc_yy;
  }
  double dL_dT[2][3];
  for (int k = 0; k < 3; k++) {
    const double a = Tm.m[0][0] * Vrk.m[k][0] + Tm.m[0][1] * Vrk.m[k][1] + Tm.m[0][2] * Vrk.m[k][2];
    const double b = Tm.m[1][0] * Vrk.m[k][0] + Tm.m[1][1] * Vrk.m[k][1] + Tm.m[1][2] * Vrk.m[k][2];
    dL_dT[0][k] = 2 * a * dL_dc_xx + b * dL_dc_xy;
    dL_dT[1][k] = 2 * b * dL_dc_yy + a * dL_dc_xy;
  }
  const double dL_dJ00 = Wm.m[0][0] * dL_dT[0][0] + Wm.m[0][1] * dL_dT[0][1] + Wm.m[0][2] * dL_dT[0][2];
  const double dL_dJ02 = Wm.m[2][0] * dL_dT[0][0] + Wm.m[2][1] * dL_dT[0][1] + Wm.m[2][2] * dL_dT[0][2];
  const double dL_dJ11 = Wm.m[1][0] * dL_dT[1][0] + Wm.m[1][1] * dL_dT[1][1] + Wm.m[1][2] * dL_dT[1][2];
  const double dL_dJ12 = Wm.m[2][0] * dL_dT[1][0] + Wm.m[2][1] * dL_dT[1][1] + Wm.m[2][2] * dL_dT[1][2];
  const double itz = 1.0 / tz, tz2 = itz * itz, tz3 = tz2 * itz;
  const double dL_dtx = x_grad_mul * -h_x * tz2 * dL_dJ02;
  const double dL_dty = y_grad_mul * -h_y * tz2 * dL_dJ12;
  const double dL_dtz = -h_x * tz2 * dL_dJ00 - h_y * tz2 * dL_dJ11 + (2 * h_x * tx) * tz3 * dL_dJ02 +
                        (2 * h_y * ty) * tz3 * dL_dJ12 - dL_dinvdepth * tz2;
  D3 gm = {view[0] * dL_dtx + view[1] * dL_dty + view[2] * dL_dtz, view[4] * dL_dtx + view[5] * dL_dty + view[6] * dL_dtz,
           view[8] * dL_dtx + view[9] * dL_dty + view[10] * dL_dtz};

  // preprocessCUDA backward (backward.cu:421-450)
  const double mw_ = proj[3] * mean.x + proj[7] * mean.y + proj[11] * mean.z + proj[15];
  const double m_w = 1.0 / (mw_ + 0.0000001);
  const double mul1 = (proj[0] * mean.x + proj[4] * mean.y + proj[8] * mean.z + proj[12]) * m_w * m_w;
  const double mul2 = (proj[1] * mean.x + proj[5] * mean.y + proj[9] * mean.z + proj[13]) * m_w * m_w;
  gm.x += (proj[0] * m_w - proj[3] * mul1) * dL_dmean2D[0] + (proj[1] * m_w - proj[3] * mul2) * dL_dmean2D[1];
  gm.y += (proj[4] * m_w - proj[7] * mul1) * dL_dmean2D[0] + (proj[5] * m_w - proj[7] * mul2) * dL_dmean2D[1];
  gm.z += (proj[8] * m_w - proj[11] * mul1) * dL_dmean2D[0] + (proj[9] * m_w - proj[11] * mul2) * dL_dmean2D[1];

  // SH backward (backward.cu:23-144)
  {
    const D3 dir_orig = {mean.x - campos[0], mean.y - campos[1], mean.z - campos[2]};
    const double len = sqrt(dir_orig.x * dir_orig.x + dir_orig.y * dir_orig.y + dir_orig.z * dir_orig.z);
    const double dx = dir_orig.x / len, dy = dir_orig.y / len, dz = dir_orig.z / len;
    const uint32_t cl = clamped[idx];
    const double g[3] = {(cl & 1u) ? 0.0 : dL_dcol[0], (cl & 2u) ? 0.0 : dL_dcol[1], (cl & 4u) ? 0.0 : dL_dcol[2]};
    const float* sh = shs + (size_t)idx * 3 * M;
    double* gsh = g_sh + (size_t)idx * 3 * M;
    double dRdx[3] = {0, 0, 0}, dRdy[3] = {0, 0, 0}, dRdz[3] = {0, 0, 0};
    for (int ch = 0; ch < 3; ch++) g_dc[3 * idx + ch] = C0 * g[ch];
#define SH(i, ch) ((double)sh[3 * (i) + (ch)])
#define GSH(i, coef) for (int ch = 0; ch < 3; ch++) gsh[3 * (i) + ch] = (coef) * g[ch];
    if (D > 0) {
      GSH(0, -C1 * dy) GSH(1, C1 * dz) GSH(2, -C1 * dx)
      for (int ch = 0; ch < 3; ch++) { dRdx[ch] = -C1 * SH(2, ch); dRdy[ch] = -C1 * SH(0, ch); dRdz[ch] = C1 * SH(1, ch); }
      if (D > 1) {
        const double xx = dx * dx, yy = dy * dy, zz = dz * dz, xy = dx * dy, yz = dy * dz, xz = dx * dz;
        GSH(3, C2[0] * xy) GSH(4, C2[1] * yz) GSH(5, C2[2] * (2.0 * zz - xx - yy)) GSH(6, C2[3] * xz) GSH(7, C2[4] * (xx - yy))
        for (int ch = 0; ch < 3; ch++) {
          dRdx[ch] += C2[0] * dy * SH(3, ch) + C2[2] * 2.0 * -dx * SH(5, ch) + C2[3] * dz * SH(6, ch) + C2[4] * 2.0 * dx * SH(7, ch);
          dRdy[ch] += C2[0] * dx * SH(3, ch) + C2[1] * dz * SH(4, ch) + C2[2] * 2.0 * -dy * SH(5, ch) + C2[4] * 2.0 * -dy * SH(7, ch);
          dRdz[ch] += C2[1] * dy * SH(4, ch) + C2[2] * 2.0 * 2.0 * dz * SH(5, ch) + C2[3] * dx * SH(6, ch);
        }
        if (D > 2) {
          GSH(8, C3[0] * dy * (3.0 * xx - yy)) GSH(9, C3[1] * xy * dz) GSH(10, C3[2] * dy * (4.0 * zz - xx - yy))
          GSH(11, C3[3] * dz * (2.0 * zz - 3.0 * xx - 3.0 * yy)) GSH(12, C3[4] * dx * (4.0 * zz - xx - yy))
          GSH(13, C3[5] * dz * (xx - yy)) GSH(14, C3[6] * dx * (xx - 3.0 * yy))
          for (int ch = 0; ch < 3; ch++) {
            dRdx[ch] += C3[0] * SH(8, ch) * 3.0 * 2.0 * xy + C3[1] * SH(9, ch) * yz + C3[2] * SH(10, ch) * -2.0 * xy +
                        C3[3] * SH(11, ch) * -3.0 * 2.0 * xz + C3[4] * SH(12, ch) * (-3.0 * xx + 4.0 * zz - yy) +
                        C3[5] * SH(13, ch) * 2.0 * xz + C3[6] * SH(14, ch) * 3.0 * (xx - yy);
            dRdy[ch] += C3[0] * SH(8, ch) * 3.0 * (xx - yy) + C3[1] * SH(9, ch) * xz + C3[2] * SH(10, ch) * (-3.0 * yy + 4.0 * zz - xx) +
                        C3[3] * SH(11, ch) * -3.0 * 2.0 * yz + C3[4] * SH(12, ch) * -2.0 * xy + C3[5] * SH(13, ch) * -2.0 * yz +
                        C3[6] * SH(14, ch) * -3.0 * 2.0 * xy;
            dRdz[ch] += C3[1] * SH(9, ch) * xy + C3[2] * SH(10, ch) * 4.0 * 2.0 * yz + C3[3] * SH(11, ch) * 3.0 * (2.0 * zz - xx - yy) +
                        C3[4] * SH(12, ch) * 4.0 * 2.0 * xz + C3[5] * SH(13, ch) * (xx - yy);
          }
        }
      }
    }
#undef SH
#undef GSH
    const double dvx = dRdx[0] * g[0] + dRdx[1] * g[1] + dRdx[2] * g[2];
    const double dvy = dRdy[0] * g[0] + dRdy[1] * g[1] + dRdy[2] * g[2];
    const double dvz = dRdz[0] * g[0] + dRdz[1] * g[1] + dRdz[2] * g[2];
    // dnormvdv (auxiliary.h:118-128)
    const double vx = dir_orig.x, vy = dir_orig.y, vz = dir_orig.z;
    const double sum2 = vx * vx + vy * vy + vz * vz;
    const double inv32 = 1.0 / sqrt(sum2 * sum2 * sum2);
    gm.x += ((sum2 - vx * vx) * dvx - vy * vx * dvy - vz * vx * dvz) * inv32;
    gm.y += (-vx * vy * dvx + (sum2 - vy * vy) * dvy - vz * vy * dvz) * inv32;
    gm.z += (-vx * vz * dvx - vy * vz * dvy + (sum2 - vz * vz) * dvz) * inv32;
  }

  // computeCov3D backward (backward.cu:354-392)
  {
    const DM3 dL_dSigma = dm3(g_cov[0], 0.5 * g_cov[1], 0.5 * g_cov[2], 0.5 * g_cov[1], g_cov[3], 0.5 * g_cov[4],
                              0.5 * g_cov[2], 0.5 * g_cov[4], g_cov[5]);
    DM3 dL_dM = mul(Mm, dL_dSigma);
    for (int c = 0; c < 3; c++)
      for (int q = 0; q < 3; q++) dL_dM.m[c][q] *= 2.0;
    const DM3 Rt = tr(R);
    DM3 dMt = tr(dL_dM);
    g_scale[3 * idx + 0] = Rt.m[0][0] * dMt.m[0][0] + Rt.m[0][1] * dMt.m[0][1] + Rt.m[0][2] * dMt.m[0][2];
    g_scale[3 * idx + 1] = Rt.m[1][0] * dMt.m[1][0] + Rt.m[1][1] * dMt.m[1][1] + Rt.m[1][2] * dMt.m[1][2];
    g_scale[3 * idx + 2] = Rt.m[2][0] * dMt.m[2][0] + Rt.m[2][1] * dMt.m[2][1] + Rt.m[2][2] * dMt.m[2][2];
    for (int q = 0; q < 3; q++) { dMt.m[0][q] *= sx; dMt.m[1][q] *= sy; dMt.m[2][q] *= sz; }
    g_rot[4 * idx + 0] = 2 * z * (dMt.m[0][1] - dMt.m[1][0]) + 2 * y * (dMt.m[2][0] - dMt.m[0][2]) + 2 * x * (dMt.m[1][2] - dMt.m[2][1]);
    g_rot[4 * idx + 1] = 2 * y * (dMt.m[1][0] + dMt.m[0][1]) + 2 * z * (dMt.m[2][0] + dMt.m[0][2]) + 2 * r * (dMt.m[1][2] - dMt.m[2][1]) - 4 * x * (dMt.m[2][2] + dMt.m[1][1]);
    g_rot[4 * idx + 2] = 2 * x * (dMt.m[1][0] + dMt.m[0][1]) + 2 * r * (dMt.m[2][0] - dMt.m[0][2]) + 2 * z * (dMt.m[1][2] + dMt.m[2][1]) - 4 * y * (dMt.m[2][2] + dMt.m[0][0]);
    g_rot[4 * idx + 3] = 2 * r * (dMt.m[0][1] - dMt.m[1][0]) + 2 * x * (dMt.m[2][0] + dMt.m[0][2]) + 2 * y * (dMt.m[1][2] + dMt.m[2][1]) - 4 * z * (dMt.m[1][1] + dMt.m[0][0]);
  }
  g_mean3D[3 * idx] = gm.x; g_mean3D[3 * idx + 1] = gm.y; g_mean3D[3 * idx + 2] = gm.z;
  double sf = 1.0;
  if (depth_threshold > 0.0) { const double q = depth / depth_threshold; sf = fmin(1.0, q * q); }   // __init__.py:171-185
  g_mean2D[2 * idx] = dL_dmean2D[0] * sf; g_mean2D[2 * idx + 1] = dL_dmean2D[1] * sf;
}

}  // namespace

extern "C" {

// accum: double[P*10], zeroed by the caller.  Returns the CUDA error code.
int truth_blend_bwd(int W, int H, const void* ranges, const void* point_list, const float* bg, const void* rec,
                    const void* n_contrib, const float* dL_dpix, const float* dL_dinvdepth, double* accum, void* stream) {
  dim3 grid((W + 15) / 16, (H + 15) / 16);
  truth_blend_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
      (const uint2*)ranges, (const uint32_t*)point_list, W, H, bg, (const Rec*)rec, (const uint32_t*)n_contrib,
      dL_dpix, dL_dinvdepth, accum);
  return (int)cudaGetLastError();
}

int truth_preprocess_bwd(int P, int D, int M, int W, int H, const float* means3D, const int* radii, const float* shs,
                         const uint8_t* clamped, const float* scales, const float* rots, double scale_modifier,
                         const float* view, const float* proj, double tan_fovx, double tan_fovy, const float* campos,
                         const double* accum, double depth_threshold, double* g_mean3D, double* g_mean2D, double* g_dc,
                         double* g_sh, double* g_opacity, double* g_scale, double* g_rot, void* stream) {
  const double h_x = W / (2.0 * tan_fovx), h_y = H / (2.0 * tan_fovy);
  truth_preprocess_bwd_kernel<<<(P + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
      P, D, M, means3D, radii, shs, clamped, scales, rots, scale_modifier, view, proj, h_x, h_y, tan_fovx, tan_fovy,
      campos, accum, depth_threshold, g_mean3D, g_mean2D, g_dc, g_sh, g_opacity, g_scale, g_rot);
  return (int)cudaGetLastError();
}

}  // extern "C"
