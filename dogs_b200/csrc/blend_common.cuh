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
// lies inside, otherwise on one of the four edges), compared with thr = ln(255*opacity) plus a
// margin that dwarfs the rounding error — so a skipped entry has alpha < 1/255 at every pixel of
// the block and skipping it cannot change any result.
__device__ __forceinline__ float quad_form(float a, float b, float c, float dx, float dy) {
  return 0.5f * (a * dx * dx + c * dy * dy) + b * dx * dy;
}
__device__ __forceinline__ bool block_reachable(const float4 co, float mx, float my, float x0,
                                                float y0, float x1, float y1, float thr) {
  const float cxp = fminf(fmaxf(mx, x0), x1), cyp = fminf(fmaxf(my, y0), y1);
  float best;
  if (cxp == mx && cyp == my) {
    best = 0.f;                                   // mean inside the rectangle
  } else {
    const float ia = __frcp_rn(co.x), ic = __frcp_rn(co.z);
    // horizontal edges: minimise over x for fixed dy
    float dy = y0 - my;
    float dx = fminf(fmaxf(mx - co.y * dy * ia, x0), x1) - mx;
    best = quad_form(co.x, co.y, co.z, dx, dy);
    dy = y1 - my;
    dx = fminf(fmaxf(mx - co.y * dy * ia, x0), x1) - mx;
    best = fminf(best, quad_form(co.x, co.y, co.z, dx, dy));
    // vertical edges: minimise over y for fixed dx
    dx = x0 - mx;
    dy = fminf(fmaxf(my - co.y * dx * ic, y0), y1) - my;
    best = fminf(best, quad_form(co.x, co.y, co.z, dx, dy));
    dx = x1 - mx;
    dy = fminf(fmaxf(my - co.y * dx * ic, y0), y1) - my;
    best = fminf(best, quad_form(co.x, co.y, co.z, dx, dy));
  }
  // !(best > ...) keeps NaNs on the safe (processed) side
  return !(best > thr + 0.01f + 1e-5f * fabsf(thr));
}

}  // namespace dogs
