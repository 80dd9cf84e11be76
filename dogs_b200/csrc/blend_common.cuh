// Helpers shared by the forward and backward blend kernels.
#pragma once
#include "common.cuh"

namespace dogs {

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// Sub-tile culling.  The reference culls (Gaussian, 16x16 tile) pairs whose best possible alpha in
// the tile is below 1/255 (rasterizer_impl.cu:52-100); every pixel of the tile still evaluates
// every surviving list entry.  Here each warp owns an 8x4 pixel block and first decides, one list
// entry per lane, whether the entry can reach alpha >= 1/255 anywhere in ITS block; only those
// entries enter the per-pixel loop.  The test is conservative — the exact minimum of the quadratic
// form 0.5*d^T C d over the block's pixel-centre rectangle (convex => attained at the mean if it
// lies inside, otherwise on an edge facing the mean), compared with thr = ln(255*opacity) plus a
// margin that dwarfs the rounding error — so a skipped entry has alpha < 1/255 at every pixel of
// the block and skipping it cannot change any result.
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float quad_form(float a, float b, float c, float dx, float dy) {
  return 0.5f * (a * dx * dx + c * dy * dy) + b * dx * dy;
}
__device__ __forceinline__ bool block_reachable(const float4 co, float mx, float my, float x0,
                                                float y0, float x1, float y1, float thr) {
  // The clamped mean is the point of the rectangle's facing edge(s).  The form grows along every ray
  // from the mean, and a ray to any point of a far edge first crosses a facing edge, so only the (at
  // most two) edges facing the mean can hold the minimum.
  const float cxp = fminf(fmaxf(mx, x0), x1), cyp = fminf(fmaxf(my, y0), y1);
  const float ex = cxp - mx, ey = cyp - my;          // 0 when the mean is inside that range
  // vertical facing edge x = cxp: minimise over y;  horizontal facing edge y = cyp: minimise over x
  // (approximate reciprocals: an error in the position of the edge minimum changes the value found there
  //  only in second order, far inside the margin below; the IEEE-rounded __frcp_rn was 45 % of this test)
  const float dyv = fminf(fmaxf(my - co.y * ex * rcp_approx(co.z), y0), y1) - my;
  const float dxh = fminf(fmaxf(mx - co.y * ey * rcp_approx(co.x), x0), x1) - mx;
  const float qv = quad_form(co.x, co.y, co.z, ex, dyv);
  const float qh = quad_form(co.x, co.y, co.z, dxh, ey);
  // mean inside the x range: only the horizontal edge counts (and vice versa); inside both: 0
  float best = (ex != 0.f) ? ((ey != 0.f) ? fminf(qv, qh) : qv) : ((ey != 0.f) ? qh : 0.f);
  // !(best > ...) keeps NaNs on the safe (processed) side
  return !(best > thr + 0.01f + 1e-5f * fabsf(thr));
}

}  // namespace dogs
