// Forward rasterizer kernels (sm_100a): per-Gaussian preprocess with the per-tile reachability
// cull folded in, instance emission in depth order, tile ranges, and the per-tile front-to-back
// alpha blend.  Arithmetic follows cuda_rasterizer/forward.cu:24-276,349-501 and
// rasterizer_impl.cu:52-220 operation by operation so that radii, keys,
// ranges, n_contrib and final_T come out bit-identical; the kernel structure does not.
#include <stdlib.h>
#include "common.cuh"

namespace dogs {

#ifndef PRE_THREADS_N
#define PRE_THREADS_N 128
#endif
constexpr int PRE_THREADS = PRE_THREADS_N;

// ---- forward.cu:119-153 ----------------------------------------------------------
__device__ __forceinline__ void cov3d_from_scale_rot(const float3 scale, float mod, const float4 rot,
                                                    float* cov3D) {
  M3 S = m3(1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f);
  S.m[0][0] = mod * scale.x;
  S.m[1][1] = mod * scale.y;
  S.m[2][2] = mod * scale.z;
  float r = rot.x, x = rot.y, y = rot.z, z = rot.w;
  M3 R = m3(1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y),
            2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x),
            2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y));
  M3 M = S * R;
  M3 Sigma = transpose(M) * M;
  cov3D[0] = Sigma.m[0][0];
  cov3D[1] = Sigma.m[0][1];
  cov3D[2] = Sigma.m[0][2];
  cov3D[3] = Sigma.m[1][1];
  cov3D[4] = Sigma.m[1][2];
  cov3D[5] = Sigma.m[2][2];
}

// ---- forward.cu:79-114 -----------------------------------------------------------
__device__ __forceinline__ float3 cov2d_from_cov3d(const float3& mean, float focal_x, float focal_y,
                                                  float tan_fovx, float tan_fovy,
                                                  const float* cov3D, const float* view) {
  float3 t = xform4x3(mean, view);
  const float limx = 1.3f * tan_fovx;
  const float limy = 1.3f * tan_fovy;
  const float txtz = t.x / t.z;
  const float tytz = t.y / t.z;
  t.x = min(limx, max(-limx, txtz)) * t.z;
  t.y = min(limy, max(-limy, tytz)) * t.z;
  M3 J = m3(focal_x / t.z, 0.0f, -(focal_x * t.x) / (t.z * t.z),
            0.0f, focal_y / t.z, -(focal_y * t.y) / (t.z * t.z),
            0, 0, 0);
  M3 Wm = m3(view[0], view[4], view[8], view[1], view[5], view[9], view[2], view[6], view[10]);
  M3 T = Wm * J;
  M3 Vrk = m3(cov3D[0], cov3D[1], cov3D[2], cov3D[1], cov3D[3], cov3D[4], cov3D[2], cov3D[4],
              cov3D[5]);
  M3 cov = transpose(T) * transpose(Vrk) * T;
  return {cov.m[0][0], cov.m[0][1], cov.m[1][1]};
}

// ---- forward.cu:24-76 ------------------------------------------------------------
// sh points at this Gaussian's M "rest" coefficients (row of 3*M floats, any address space).
__device__ __forceinline__ V3 color_from_sh(int deg, V3 pos, V3 campos, V3 dc, const float* sh,
                                            uint32_t& clamped) {
  V3 dir = pos - campos;
  dir = dir / sqrtf(dot(dir, dir));
  V3 result = SH_C0 * dc;
#define SHV(i) (V3{sh[3 * (i)], sh[3 * (i) + 1], sh[3 * (i) + 2]})
  if (deg > 0) {
    float x = dir.x, y = dir.y, z = dir.z;
    result = result - SH_C1 * y * SHV(0) + SH_C1 * z * SHV(1) - SH_C1 * x * SHV(2);
    if (deg > 1) {
      float xx = x * x, yy = y * y, zz = z * z;
      float xy = x * y, yz = y * z, xz = x * z;
      result = result + SH_C2_0 * xy * SHV(3) + SH_C2_1 * yz * SHV(4) +
               SH_C2_2 * (2.0f * zz - xx - yy) * SHV(5) + SH_C2_3 * xz * SHV(6) +
               SH_C2_4 * (xx - yy) * SHV(7);
      if (deg > 2) {
        result = result + SH_C3_0 * y * (3.0f * xx - yy) * SHV(8) + SH_C3_1 * xy * z * SHV(9) +
                 SH_C3_2 * y * (4.0f * zz - xx - yy) * SHV(10) +
                 SH_C3_3 * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * SHV(11) +
                 SH_C3_4 * x * (4.0f * zz - xx - yy) * SHV(12) + SH_C3_5 * z * (xx - yy) * SHV(13) +
                 SH_C3_6 * x * (xx - 3.0f * yy) * SHV(14);
      }
    }
  }
#undef SHV
  result.x += 0.5f;
  result.y += 0.5f;
  result.z += 0.5f;
  clamped = (result.x < 0 ? 1u : 0u) | (result.y < 0 ? 2u : 0u) | (result.z < 0 ? 4u : 0u);
  return {fmaxf(result.x, 0.0f), fmaxf(result.y, 0.0f), fmaxf(result.z, 0.0f)};
}

__device__ __forceinline__ float emit_rcp(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// One thread per Gaussian.  The CTA's slice of the SH tensor (the bulk of the input bytes:
// 180 of 236 B per Gaussian) is contiguous, so it is staged into shared memory with coalesced
// 16-byte loads and read back row-wise at a stride of 3*M floats (odd => conflict-free).
// Replaces preprocessCUDA (forward.cu:156-276) and the counting half of duplicateWithKeys
// (rasterizer_impl.cu:120-190): tiles[idx] is the number of instances that survive the cull.
__global__ void __launch_bounds__(PRE_THREADS)
preprocess_fwd_kernel(int P, int D, int M, const float* __restrict__ means3D,
                      const float* __restrict__ scales, float scale_modifier,
                      const float* __restrict__ rotations, const float* __restrict__ opacities,
                      const float* __restrict__ dc, const float* __restrict__ shs,
                      const float* __restrict__ cov3D_precomp,
                      const float* __restrict__ colors_precomp, const float* __restrict__ view,
                      const float* __restrict__ proj, const float* __restrict__ cam_pos, int W,
                      int H, float tan_fovx, float tan_fovy, float focal_x, float focal_y,
                      int* __restrict__ radii, SplatRecord* __restrict__ rec,
                      uint8_t* __restrict__ clamped_out, uint32_t* __restrict__ tiles,
                      uint32_t* __restrict__ depth_key, uint4* __restrict__ emit_rec,
                      uint32_t* __restrict__ ctl,
                      uint32_t gx, uint32_t gy, bool prefiltered, bool antialiasing, bool raw_inputs,
                      uint32_t walk_off) {
  extern __shared__ __align__(128) float4 sh_stage4[];
  float* sh_stage = reinterpret_cast<float*>(sh_stage4);
  __shared__ __align__(8) uint64_t sh_bar;
  const int first = blockIdx.x * PRE_THREADS;
  const int idx = first + threadIdx.x;
  const bool use_sh = (colors_precomp == nullptr) && D > 0 && M > 0;
  // The CTA's SH slice (count * 3M floats, contiguous) goes to shared memory as ONE bulk copy issued by one
  // thread; it lands while the threads project their Gaussians and build the covariances, and is awaited
  // (mbarrier) only where the colours are evaluated.  Fallback for a slice that is not a multiple of 16 bytes
  // (ragged last CTA) or a misaligned tensor: coalesced loads through the LSU.
  bool bulk = false;
  if (use_sh) {
    const int count = min(PRE_THREADS, P - first);
    const size_t nfl = (size_t)count * 3 * M;
    const float* src = shs + (size_t)first * 3 * M;
    bulk = (((uintptr_t)src) & 15) == 0 && ((nfl * sizeof(float)) & 15) == 0;
    if (bulk) {
      if (threadIdx.x == 0) {
        mbar_init(&sh_bar, 1);
        bulk_load(sh_stage, src, (uint32_t)(nfl * sizeof(float)), &sh_bar);
      }
    } else {
      for (size_t i = threadIdx.x; i < nfl; i += PRE_THREADS) sh_stage[i] = __ldg(src + i);
    }
    __syncthreads();     // the barrier is initialised (bulk) / the slice is staged (fallback)
  }
  // (no early return: the whole warp meets again for the cooperative rect walk below)

  int radius_out = 0;
  uint32_t ntiles = 0, dkey = 0xffffffffu;
  unsigned long long mask = 0ull;
  uint32_t rect_word = 0;
  // the rect of this lane's Gaussian, walked by the whole warp below (w_n = 0: culled, nothing to walk)
  float4 w_co = {0.f, 0.f, 0.f, 0.f};
  float2 w_xy = {0.f, 0.f};
  float w_thr = 0.f;
  int w_x0 = 0, w_y0 = 0, w_w = 0;
  uint32_t w_n = 0;
  do {
    if (idx >= P) break;
    float3 p_orig = {means3D[3 * idx], means3D[3 * idx + 1], means3D[3 * idx + 2]};
    // The opacity and the dc colour are requested here, with the position, not where they are used (after the
    // frustum and rect tests): loads are not moved across those branches, and each was a memory round trip of its
    // own.  (Hoisting the scale / rotation loads as well changes ptxas' FMA contraction in the covariance code —
    // conics 1-2 ulp off the reference's — so those stay where they are.)
    float op_in;
    float3 dc_in = {0.f, 0.f, 0.f};
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(op_in) : "l"(opacities + idx));
    if (colors_precomp == nullptr) {
      asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(dc_in.x) : "l"(dc + 3 * idx));
      asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(dc_in.y) : "l"(dc + 3 * idx + 1));
      asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(dc_in.z) : "l"(dc + 3 * idx + 2));
    }
    // auxiliary.h:150-175 in_frustum
    float4 p_hom = xform4x4(p_orig, proj);
    float p_w = 1.0f / (p_hom.w + 0.0000001f);
    float3 p_proj = {p_hom.x * p_w, p_hom.y * p_w, p_hom.z * p_w};
    float3 p_view = xform4x3(p_orig, view);
    if (p_view.z <= 0.2f) {
      if (prefiltered) {
        printf("Point is filtered although prefiltered is set. This shouldn't happen!");
        __trap();
      }
      break;
    }
    float cov3D[6];
    if (cov3D_precomp != nullptr) {
#pragma unroll
      for (int i = 0; i < 6; i++) cov3D[i] = cov3D_precomp[6 * idx + i];
    } else {
      float3 sc = {scales[3 * idx], scales[3 * idx + 1], scales[3 * idx + 2]};
      float4 q = reinterpret_cast<const float4*>(rotations)[idx];
      if (raw_inputs) { sc = act_exp3(sc); q = act_normalize(q); }
      cov3d_from_scale_rot(sc, scale_modifier, q, cov3D);
    }
    float3 cov = cov2d_from_cov3d(p_orig, focal_x, focal_y, tan_fovx, tan_fovy, cov3D, view);

    constexpr float h_var = 0.3f;
    const float det_cov = cov.x * cov.z - cov.y * cov.y;
    cov.x += h_var;
    cov.z += h_var;
    const float det_cov_plus_h_cov = cov.x * cov.z - cov.y * cov.y;
    float h_convolution_scaling = 1.0f;
    if (antialiasing) h_convolution_scaling = sqrt(max(0.000025f, det_cov / det_cov_plus_h_cov));
    const float det = det_cov_plus_h_cov;
    if (det == 0.0f) break;
    float det_inv = 1.f / det;
    float3 conic = {cov.z * det_inv, -cov.y * det_inv, cov.x * det_inv};

    float mid = 0.5f * (cov.x + cov.z);
    float lambda1 = mid + sqrt(max(0.1f, mid * mid - det));
    float lambda2 = mid - sqrt(max(0.1f, mid * mid - det));
    float my_radius = ceil(3.f * sqrt(max(lambda1, lambda2)));
    float2 point_image = {ndc2Pix(p_proj.x, W), ndc2Pix(p_proj.y, H)};
    uint2 rect_min, rect_max;
    get_rect(point_image, (int)my_radius, rect_min, rect_max, gx, gy);
    if ((rect_max.x - rect_min.x) * (rect_max.y - rect_min.y) == 0) break;

    float3 rgb;
    uint32_t clamped = 0;
    if (colors_precomp == nullptr) {
      V3 pos = {p_orig.x, p_orig.y, p_orig.z};
      V3 cp = {cam_pos[0], cam_pos[1], cam_pos[2]};
      V3 d0 = {dc_in.x, dc_in.y, dc_in.z};
      if (bulk) mbar_wait(&sh_bar, 0);
      V3 c = color_from_sh(D, pos, cp, d0, sh_stage + (size_t)threadIdx.x * 3 * M, clamped);
      rgb = {c.x, c.y, c.z};
    } else {
      rgb = {colors_precomp[3 * idx], colors_precomp[3 * idx + 1], colors_precomp[3 * idx + 2]};
    }
    clamped_out[idx] = (uint8_t)clamped;

    const float depth = p_view.z;
    const float opacity = (raw_inputs ? act_sigmoid(op_in) : op_in) * h_convolution_scaling;
    radius_out = (int)my_radius;
    const float4 co = {conic.x, conic.y, conic.z, opacity};
    SplatRecord r;
    r.a = co;
    r.b = {point_image.x, point_image.y, 1.f / depth, rgb.x};
    // rasterizer_impl.cu:151-177: count the tiles of the rect the Gaussian can actually reach.
    const float thr = logf(co.w / (1.0f / 255.0f));
    r.c = {rgb.y, rgb.z, depth, thr};
    rec[idx] = r;

    const int rw = rect_max.x - rect_min.x;
    const bool small = rw * (int)(rect_max.y - rect_min.y) <= 64 && rect_min.x < 2048u && rect_min.y < 2048u;
    rect_word = small ? (rect_min.x | (rect_min.y << 11) | ((uint32_t)rw << 22)) : 0u;   // mask fits 64 bits
    w_co = co; w_xy = point_image; w_thr = thr;
    w_x0 = rect_min.x; w_y0 = rect_min.y; w_w = rw; w_n = (uint32_t)(rw * (int)(rect_max.y - rect_min.y));
    dkey = __float_as_uint(depth);
  } while (0);
  // Reachability walk over the rects (rasterizer_impl.cu:151-177), by the whole warp over the FLATTENED list of the
  // (Gaussian, tile) pairs of its 32 Gaussians, 32 pairs per step.  (Second version.  In the first one every thread
  // walked its own rect, and only rects of more than 16 tiles — the largest ~1000 tiles — were handed to the warp:
  // the walk was half of this kernel's instructions at 12 of 32 lanes active, because rect sizes within a warp
  // differ by an order of magnitude.)  A lane finds the owner of its pair by binary search over the lanes'
  // offsets, reads the owner's parameters from shared memory, tests one tile; the owners collect their bits of
  // the step's ballot.  The parameter slots alias the warp's own rows of the SH slice, which its lanes have
  // finished reading (walk_off = 0xffffffff), or lie behind the slice (small SH degrees / precomputed colours).
  {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4* slot = walk_off == 0xffffffffu
                       ? reinterpret_cast<float4*>(reinterpret_cast<char*>(sh_stage) + (size_t)warp * 32 * 3 * M * sizeof(float))
                       : reinterpret_cast<float4*>(reinterpret_cast<char*>(sh_stage) + walk_off) + warp * 32 * 3;
    uint32_t incl = w_n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(full, incl, o);
      if (lane >= o) incl += y;
    }
    const uint32_t woff = incl - w_n, total = __shfl_sync(full, incl, 31);
    __syncwarp();                      // every lane of the warp is done with its SH row
    if (w_n) {
      const float2 rc = tile_rcps(w_co);
      slot[3 * lane] = w_co;
      slot[3 * lane + 1] = make_float4(w_xy.x, w_xy.y, w_thr, rc.x);
      slot[3 * lane + 2] = make_float4(rc.y, __int_as_float(w_x0), __int_as_float(w_y0), __int_as_float(w_w));
    }
    __syncwarp();
    uint32_t cnt = 0, mlo = 0, mhi = 0;
    for (uint32_t base = 0; base < total; base += 32) {
      const uint32_t t = base + lane;
      int g = 0;                       // largest lane with woff <= t: the owner of pair t
#pragma unroll
      for (int step = 16; step >= 1; step >>= 1) {
        const uint32_t o = __shfl_sync(full, woff, g + step);
        if (o <= t) g += step;
      }
      const uint32_t k = t - __shfl_sync(full, woff, g);
      bool r = false;
      if (t < total) {
        const float4 a = slot[3 * g], b = slot[3 * g + 1], c = slot[3 * g + 2];
        const int w = __float_as_int(c.w);
        // k / w: (k + 0.5) / w stays 1/(2w) away from every integer, so for these ranges the truncated product
        // with an approximate reciprocal is the exact quotient; huge rects take the integer division
        const int ty = (w <= 128 && k < 16384u) ? (int)(((float)k + 0.5f) * emit_rcp((float)w)) : (int)(k / (uint32_t)w);
        const int tx = (int)k - ty * w;
        r = tile_reachable(a, make_float2(b.x, b.y), __float_as_int(c.y) + tx, __float_as_int(c.z) + ty, b.z,
                           make_float2(b.w, c.x));
      }
      const unsigned bal = __ballot_sync(full, r);
      // this lane's span [sa, sb) of the step's 32 pairs
      const int sa = max((int)woff - (int)base, 0), sb = min((int)(woff + w_n) - (int)base, 32);
      if (sb > sa) {
        const uint32_t seg = (bal >> sa) & (sb - sa == 32 ? 0xffffffffu : (1u << (sb - sa)) - 1u);
        cnt += __popc(seg);
        const int j0 = (int)base + sa - (int)woff;      // rect-local index of the span's first pair
        if (j0 < 64) {                                   // (the mask is only used for rects of <= 64 tiles)
          const unsigned long long m = (unsigned long long)seg << j0;
          mlo |= (uint32_t)m;
          mhi |= (uint32_t)(m >> 32);
        }
      }
    }
    ntiles = cnt;
    mask = ((unsigned long long)mhi << 32) | mlo;
  }
  // range of the depth keys that take part in the sort (ctl[0] = min, ctl[1] = max): one reduction per CTA
  {
    __shared__ uint32_t s_min[PRE_THREADS / 32], s_max[PRE_THREADS / 32];
    uint32_t kmin = (idx < P && ntiles) ? dkey : 0xffffffffu, kmax = (idx < P && ntiles) ? dkey : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
      kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
    }
    if ((threadIdx.x & 31) == 0) { s_min[threadIdx.x >> 5] = kmin; s_max[threadIdx.x >> 5] = kmax; }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
      for (int w = 1; w < PRE_THREADS / 32; w++) { kmin = min(kmin, s_min[w]); kmax = max(kmax, s_max[w]); }
      if (kmin <= kmax) { atomicMin(&ctl[0], kmin); atomicMax(&ctl[1], kmax); }
    }
  }
  if (bulk) mbar_wait(&sh_bar, 0);     // nobody leaves while the copy may still be writing this CTA's shared memory
  if (idx >= P) return;
  radii[idx] = radius_out;
  tiles[idx] = ntiles;
  emit_rec[idx] = make_uint4(rect_word, ntiles, (uint32_t)mask, (uint32_t)(mask >> 32));
  depth_key[idx] = ntiles ? dkey : 0xffffffffu;
}

// The depth keys as preprocess wrote them (the sort ping-pongs through that array): rebuilt from the records
// when the depth stage has to be repeated on more key bits.
__global__ void __launch_bounds__(256)
depth_keys_kernel(int P, const SplatRecord* __restrict__ rec, const uint32_t* __restrict__ tiles,
                  uint32_t* __restrict__ depth_key) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  depth_key[i] = tiles[i] ? __float_as_uint(rec[i].c.z) : 0xffffffffu;
}

// emit_rec -> depth order (one 16-byte gather per Gaussian); .y becomes the Gaussian id.  The CTA also leaves the sum
// of its 256 instance counts in block_sum[blockIdx.x]: scanned by one CTA (radix_rowscan_kernel) these are the
// bases from which emit_instances_kernel derives every offset itself, so no per-Gaussian offset array is written
// or read (the three-kernel scan over P counts this replaces: 0.031 ms at 3 M Gaussians).
__global__ void __launch_bounds__(256)
gather_emit_kernel(int P, const uint32_t* __restrict__ sorted_idx, const uint4* __restrict__ emit_rec,
                   uint4* __restrict__ sorted_emit, uint32_t* __restrict__ sorted_cnt, uint32_t* __restrict__ block_sum) {
  pdl_wait();
  pdl_trigger();
  __shared__ uint32_t wsum[8];
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t c = 0;
  if (j < P) {
    const uint32_t id = sorted_idx[j];
    const uint4 e = emit_rec[id];
    c = e.y;
    sorted_cnt[j] = c;
    sorted_emit[j] = make_uint4(e.x, id, e.z, e.w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) t += wsum[w];
    block_sum[blockIdx.x] = t;
  }
}

// Small models: the gather and the exclusive scan of the instance counts in depth order (duplicateWithKeys'
// offsets, rasterizer_impl.cu:447-452 + cub::DeviceScan::InclusiveSum) in ONE kernel instead of four.  The scan is
// chained: a CTA takes its chunk number from a ticket counter, publishes the chunk total, and warp 0 looks back
// over the status words of the chunks before it, 32 at a time (2 flag bits + 30-bit value: total first, inclusive
// prefix once known).  A chunk a CTA waits for is owned by a CTA that already runs, so the wait cannot deadlock.
// (Measured: 0.026 vs 0.025 ms at C1 with three launches less; at 3 M Gaussians the chain of prefixes sets the
// pace — 0.069 ms against 0.055 ms for gather + three-kernel scan, and wider look-back windows, back-off and larger
// chunks all made it slower — so large models keep the separate kernels.)
constexpr int GS_THREADS = 256, GS_ITEMS = 4, GS_CHUNK = GS_THREADS * GS_ITEMS;
constexpr uint32_t GS_LOCAL = 1u << 30, GS_INCL = 2u << 30, GS_VALUE = (1u << 30) - 1u;
constexpr size_t GS_CHAINED_MAX = 1u << 18;     // Gaussians up to which the chained kernel is used
__global__ void __launch_bounds__(GS_THREADS)
gather_scan_kernel(int P, const uint32_t* __restrict__ sorted_idx, const uint4* __restrict__ emit_rec,
                   uint4* __restrict__ sorted_emit, uint32_t* __restrict__ sorted_cnt, uint32_t* __restrict__ offsets,
                   uint32_t* __restrict__ status, uint32_t* __restrict__ ticket, uint32_t* __restrict__ total_out) {
  pdl_wait();
  pdl_trigger();
  __shared__ uint32_t wsum[GS_THREADS / 32];
  __shared__ uint32_t s_chunk, s_base;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_chunk = atomicAdd(ticket, 1u);
  __syncthreads();
  const uint32_t chunk = s_chunk;
  const int j0 = (int)chunk * GS_CHUNK + tid * GS_ITEMS;
  uint32_t id[GS_ITEMS];
  uint4 e[GS_ITEMS];
#pragma unroll
  for (int k = 0; k < GS_ITEMS; k++) id[k] = j0 + k < P ? sorted_idx[j0 + k] : 0u;
#pragma unroll
  for (int k = 0; k < GS_ITEMS; k++) e[k] = j0 + k < P ? emit_rec[id[k]] : make_uint4(0u, 0u, 0u, 0u);
  uint32_t sum = 0;
#pragma unroll
  for (int k = 0; k < GS_ITEMS; k++) sum += e[k].y;
  uint32_t x = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) wsum[warp] = x;
  __syncthreads();
  uint32_t woff = 0;
#pragma unroll
  for (int w = 0; w < GS_THREADS / 32; w++) if (w < warp) woff += wsum[w];
  if (warp == 0) {
    uint32_t total = lane < GS_THREADS / 32 ? wsum[lane] : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
    uint32_t before = 0;
    if (chunk > 0) {
      if (lane == 0) asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(status + chunk), "r"(GS_LOCAL | total) : "memory");
      int t = (int)chunk - 1;                    // lane l looks at chunk t - l
      while (true) {
        uint32_t sv = GS_INCL;                    // before chunk 0: an inclusive prefix of 0
        if (t - lane >= 0) asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(sv) : "l"(status + (t - lane)) : "memory");
        const uint32_t f = sv >> 30;
        const uint32_t incl = __ballot_sync(0xffffffffu, f == 2), ready = __ballot_sync(0xffffffffu, f != 0);
        // lanes 0..last are needed: up to the nearest inclusive prefix, or all 32
        const int last = incl ? __ffs(incl) - 1 : 31;
        const uint32_t need = last == 31 ? 0xffffffffu : (2u << last) - 1u;
        if ((ready & need) != need) continue;     // something in that span is not published yet: poll again
        uint32_t v = lane <= last ? (sv & GS_VALUE) : 0u;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        before += v;
        if (incl) break;
        t -= 32;
      }
    }
    if (lane == 0) {
      asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(status + chunk), "r"(GS_INCL | (before + total)) : "memory");
      s_base = before;
      if ((int)(chunk + 1) * GS_CHUNK >= P) *total_out = before + total;   // the last chunk: R
    }
  }
  __syncthreads();
  uint32_t off = s_base + woff + x - sum;
#pragma unroll
  for (int k = 0; k < GS_ITEMS; k++) {
    if (j0 + k < P) {
      sorted_cnt[j0 + k] = e[k].y;
      offsets[j0 + k] = off;
      sorted_emit[j0 + k] = make_uint4(e[k].x, id[k], e[k].z, e[k].w);
    }
    off += e[k].y;
  }
}

// Position of the (k+1)-th set bit of a 64-bit mask (k < popcount).
__device__ __forceinline__ int kth_set_bit64(uint32_t lo, uint32_t hi, int k) {
  const int nlo = __popc(lo);
  uint32_t w = lo;
  int base = 0;
  if (k >= nlo) { w = hi; k -= nlo; base = 32; }
  // binary search with popcounts over halves of the word
  int pos = 0;
#pragma unroll
  for (int width = 16; width >= 1; width >>= 1) {
    const uint32_t low = w & ((1u << width) - 1u);
    const int c = __popc(low);
    if (k >= c) { k -= c; w >>= width; pos += width; } else { w = low; }
  }
  return base + pos;
}

// Write (tile id, Gaussian id) for every reachable tile of every Gaussian, in depth order (the
// emitting half of duplicateWithKeys, rasterizer_impl.cu:151-177).  Instances are therefore already
// in (depth, Gaussian id) order; the stable sort by tile id that follows reproduces the order of
// the reference's 64-bit (tile|depth) sort without ever materialising 64-bit keys.
// A warp owns 32 consecutive depth-sorted Gaussians and therefore one contiguous run of the output.
// The run is written 32 slots at a time, one slot per lane (full 128-byte stores): the lane finds the
// owner of its slot by binary search over the lanes' offsets (shuffles) and takes the k-th set bit of
// the owner's reachable-tile mask recorded by preprocess.  (The first version gave each thread its own
// Gaussian and a divergent loop over its mask: ncu thread efficiency 12/32, 42 % of DRAM peak.)
// Gaussians whose rect exceeds 64 tiles have no mask; the warp re-runs the test for them afterwards.
__global__ void __launch_bounds__(256)
emit_instances_kernel(int P, const uint4* __restrict__ sorted_emit, const uint32_t* __restrict__ sorted_cnt,
                      const uint32_t* __restrict__ offsets, const uint32_t* __restrict__ block_base,
                      const SplatRecord* __restrict__ rec,
                      const int* __restrict__ radii, uint32_t* __restrict__ tile_keys,
                      uint32_t* __restrict__ vals, uint32_t gx, uint32_t gy, uint32_t capacity) {
  pdl_wait();
  pdl_trigger();
  const unsigned full = 0xffffffffu;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const uint32_t cnt = i < P ? sorted_cnt[i] : 0u;
  uint4 e = make_uint4(0u, 0u, 0u, 0u);
  if (cnt) e = sorted_emit[i];
  // off = exclusive scan of sorted_cnt: non-decreasing, equal across zero-count entries.  Large models: the CTA's
  // base (scanned block sums of gather_emit_kernel) + a scan over the CTA's own 256 counts; small models: the
  // offsets the chained gather + scan kernel wrote.
  uint32_t off;
  if (block_base) {
    __shared__ uint32_t wsum[8];
    uint32_t x = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(full, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) wsum[threadIdx.x >> 5] = x;
    __syncthreads();
    uint32_t woff = block_base[blockIdx.x];
#pragma unroll
    for (int w = 0; w < 8; w++) if (w < (int)(threadIdx.x >> 5)) woff += wsum[w];
    off = i < P ? woff + x - cnt : 0xffffffffu;
  } else {
    off = i < P ? offsets[i] : 0xffffffffu;
  }
  const uint32_t begin = __shfl_sync(full, off, 0);
  uint32_t end = off + cnt;                       // of the last in-range lane
  {
    const int last_lane = max(0, min(31, P - 1 - (i - lane)));
    end = __shfl_sync(full, end, last_lane);
  }
  const uint32_t rw = e.x >> 22;
  for (uint32_t base = begin; base < end; base += 32) {
    const uint32_t t = base + lane;
    int g = 0;                                    // largest lane with off <= t
#pragma unroll
    for (int step = 16; step >= 1; step >>= 1) {
      const uint32_t o = __shfl_sync(full, off, g + step);
      if (o <= t) g += step;
    }
    const uint32_t g_off = __shfl_sync(full, off, g);
    const uint32_t g_x = __shfl_sync(full, e.x, g), g_id = __shfl_sync(full, e.y, g);
    const uint32_t g_lo = __shfl_sync(full, e.z, g), g_hi = __shfl_sync(full, e.w, g);
    const uint32_t g_rw = g_x >> 22;
    if (t < end && t < capacity && g_rw) {
      const int bit = kth_set_bit64(g_lo, g_hi, (int)(t - g_off));
      const uint32_t x0 = g_x & 2047u, y0 = (g_x >> 11) & 2047u;
      // bit < 64 and rect width <= 64: (bit + 0.5) / width stays 1/128 away from every integer, far beyond the
      // error of the approximate reciprocal, so the truncation is the exact quotient (the integer division was
      // ~20 instructions of the ~100 per 32 instances)
      const uint32_t ty = (uint32_t)(((float)bit + 0.5f) * emit_rcp((float)g_rw)), tx = (uint32_t)bit - ty * g_rw;
      tile_keys[t] = (y0 + ty) * gx + x0 + tx;
      vals[t] = g_id;
    }
  }
  // rects of more than 64 tiles: no mask; the warp re-runs the reachability test, 32 tiles per step
  unsigned big = __ballot_sync(full, cnt != 0 && rw == 0);
  while (big) {
    const int src = __ffs(big) - 1;
    big &= big - 1;
    const uint32_t idx = __shfl_sync(full, e.y, src);
    uint32_t o = __shfl_sync(full, off, src);
    // the count was taken by preprocess with the same test; should the two ever disagree (different FMA placement
    // at the two inlining sites), stay inside this Gaussian's own slots
    const uint32_t o_end = o + __shfl_sync(full, cnt, src);
    const float4 b = rec[idx].b;
    const float2 xy = {b.x, b.y};
    uint2 rect_min, rect_max;
    get_rect(xy, radii[idx], rect_min, rect_max, gx, gy);
    const float4 co = rec[idx].a;
    const float thr = logf(co.w / (1.0f / 255.0f));
    const int w = rect_max.x - rect_min.x, n = w * (int)(rect_max.y - rect_min.y);
    const float2 rcps = tile_rcps(co);
    for (int base = 0; base < n; base += 32) {
      const int t = base + lane;
      bool r = false;
      int x = 0, y = 0;
      if (t < n) {
        const int ty = t / w;
        x = rect_min.x + (t - ty * w);
        y = rect_min.y + ty;
        r = tile_reachable(co, xy, x, y, thr, rcps);
      }
      const unsigned bal = __ballot_sync(full, r);
      if (r) {
        const uint32_t pos = o + __popc(bal & ((1u << lane) - 1u));
        if (pos < capacity && pos < o_end) {
          tile_keys[pos] = y * gx + x;
          vals[pos] = idx;
        }
      }
      o += __popc(bal);
    }
  }
}

// identifyTileRanges (rasterizer_impl.cu:195-220) on 32-bit tile keys; no pad entries exist.
// Eight consecutive keys per thread (two 16-byte loads + the predecessor, all requested before the first compare).
constexpr int TR_KEYS = 8;
__global__ void __launch_bounds__(256)
tile_ranges_kernel(uint32_t L_cap, const uint32_t* __restrict__ L_dev, const uint32_t* __restrict__ tile_keys,
                   uint2* __restrict__ ranges, uint32_t T) {
  pdl_wait();
  pdl_trigger();
  const uint32_t L = L_dev ? min(*L_dev, L_cap) : L_cap;
  const uint32_t i0 = (blockIdx.x * blockDim.x + threadIdx.x) * (uint32_t)TR_KEYS;
  if (i0 >= L) return;
  const uint32_t n = min((uint32_t)TR_KEYS, L - i0);
  uint32_t k[TR_KEYS];
#pragma unroll
  for (int q = 0; q < TR_KEYS; q++) k[q] = 0u;
  uint32_t prev = i0 ? tile_keys[i0 - 1] : 0u;
  if (n == TR_KEYS) {
    const uint4 v = *reinterpret_cast<const uint4*>(tile_keys + i0), w = *reinterpret_cast<const uint4*>(tile_keys + i0 + 4);
    k[0] = v.x; k[1] = v.y; k[2] = v.z; k[3] = v.w; k[4] = w.x; k[5] = w.y; k[6] = w.z; k[7] = w.w;
  } else {
    for (uint32_t q = 0; q < n; q++) k[q] = tile_keys[i0 + q];
  }
#pragma unroll
  for (uint32_t q = 0; q < (uint32_t)TR_KEYS; q++) {
    if (q >= n) break;
    const uint32_t idx = i0 + q, cur = k[q];
    if (cur < T) {        // (a slot the emit stage left unwritten would carry garbage: never index with it)
      if (idx == 0) {
        ranges[cur].x = 0;
      } else if (cur != prev) {
        if (prev < T) ranges[prev].y = idx;
        ranges[cur].x = idx;
      }
      if (idx == L - 1) ranges[cur].y = L;
    }
    prev = cur;
  }
}

// Launch order of the per-tile blend kernels: tile ids by list length, longest first.  List lengths are
// heavy-tailed (a few tiles hold 5-10x the mean), and with one CTA per tile in row-major order the long ones
// that happen to come last leave most SMs idle while they finish.  One CTA bins the T lengths into 1024 buckets
// (approximate order is all the scheduler needs) and writes the permutation.
__global__ void __launch_bounds__(1024)
tile_order_kernel(uint32_t T, const uint2* __restrict__ ranges, uint32_t* __restrict__ order) {
  __shared__ uint32_t bins[1024];
  __shared__ uint32_t wsum[32];
  __shared__ uint32_t smax;
  const uint32_t tid = threadIdx.x;
  bins[tid] = 0;
  if (tid == 0) smax = 0;
  pdl_wait();
  pdl_trigger();
  __syncthreads();
  uint32_t m = 0;
  for (uint32_t t = tid; t < T; t += 1024) m = max(m, ranges[t].y - ranges[t].x);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((tid & 31) == 0) atomicMax(&smax, m);
  __syncthreads();
  const float scale = 1023.0f / (float)max(smax, 1u);
  for (uint32_t t = tid; t < T; t += 1024) {
    const uint32_t len = ranges[t].y - ranges[t].x;
    atomicAdd(&bins[1023u - min(1023u, (uint32_t)((float)len * scale))], 1u);
  }
  __syncthreads();
  // exclusive scan of the 1024 bins (one per thread)
  const uint32_t v = bins[tid];
  uint32_t x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if ((tid & 31) >= (uint32_t)o) x += y;
  }
  if ((tid & 31) == 31) wsum[tid >> 5] = x;
  __syncthreads();
  if (tid < 32) {
    uint32_t w = wsum[tid], y2 = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, y2, o);
      if (tid >= (uint32_t)o) y2 += y;
    }
    wsum[tid] = y2 - w;
  }
  __syncthreads();
  bins[tid] = wsum[tid >> 5] + x - v;
  __syncthreads();
  for (uint32_t t = tid; t < T; t += 1024) {
    const uint32_t len = ranges[t].y - ranges[t].x;
    order[atomicAdd(&bins[1023u - min(1023u, (uint32_t)((float)len * scale))], 1u)] = t;
  }
}

// ---- visible_filter / mark_visible ---------------------------------------------------
// filter_preprocessCUDA (forward.cu:279-344): radii only, without the +0.3 low-pass.
__global__ void __launch_bounds__(256)
filter_kernel(int P, const float* __restrict__ means3D, const float* __restrict__ scales,
              float scale_modifier, const float* __restrict__ rotations,
              const float* __restrict__ cov3D_precomp, const float* __restrict__ view,
              const float* __restrict__ proj, int W, int H, float tan_fovx, float tan_fovy,
              float focal_x, float focal_y, int* __restrict__ radii, uint32_t gx, uint32_t gy,
              bool prefiltered, bool raw_inputs) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= P) return;
  int radius_out = 0;
  do {
    float3 p_orig = {means3D[3 * idx], means3D[3 * idx + 1], means3D[3 * idx + 2]};
    float4 p_hom = xform4x4(p_orig, proj);
    float p_w = 1.0f / (p_hom.w + 0.0000001f);
    float3 p_proj = {p_hom.x * p_w, p_hom.y * p_w, p_hom.z * p_w};
    float3 p_view = xform4x3(p_orig, view);
    if (p_view.z <= 0.2f) {
      if (prefiltered) {
        printf("Point is filtered although prefiltered is set. This shouldn't happen!");
        __trap();
      }
      break;
    }
    float cov3D[6];
    if (cov3D_precomp != nullptr) {
#pragma unroll
      for (int i = 0; i < 6; i++) cov3D[i] = cov3D_precomp[6 * idx + i];
    } else {
      float3 sc = {scales[3 * idx], scales[3 * idx + 1], scales[3 * idx + 2]};
      float4 q = reinterpret_cast<const float4*>(rotations)[idx];
      if (raw_inputs) { sc = act_exp3(sc); q = act_normalize(q); }
      cov3d_from_scale_rot(sc, scale_modifier, q, cov3D);
    }
    float3 cov = cov2d_from_cov3d(p_orig, focal_x, focal_y, tan_fovx, tan_fovy, cov3D, view);
    float det = (cov.x * cov.z - cov.y * cov.y);
    if (det == 0.0f) break;
    float mid = 0.5f * (cov.x + cov.z);
    float lambda1 = mid + sqrt(max(0.1f, mid * mid - det));
    float lambda2 = mid - sqrt(max(0.1f, mid * mid - det));
    float my_radius = ceil(3.f * sqrt(max(lambda1, lambda2)));
    float2 point_image = {ndc2Pix(p_proj.x, W), ndc2Pix(p_proj.y, H)};
    uint2 rect_min, rect_max;
    get_rect(point_image, (int)my_radius, rect_min, rect_max, gx, gy);
    if ((rect_max.x - rect_min.x) * (rect_max.y - rect_min.y) == 0) break;
    radius_out = (int)my_radius;
  } while (0);
  radii[idx] = radius_out;
}

// checkFrustum (rasterizer_impl.cu:104-116)
__global__ void __launch_bounds__(256)
mark_visible_kernel(int P, const float* __restrict__ means3D, const float* __restrict__ view,
                    uint8_t* __restrict__ present) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= P) return;
  float3 p = {means3D[3 * idx], means3D[3 * idx + 1], means3D[3 * idx + 2]};
  float3 p_view = xform4x3(p, view);
  present[idx] = p_view.z > 0.2f ? 1 : 0;
}

// ---- host-side stage launchers ---------------------------------------------------------
int launch_preprocess_fwd(const DgrParams& prm, const float* means3D, const float* dc,
                          const float* sh, const float* colors_precomp, const float* opacities,
                          const float* scales, const float* rotations, const float* cov3D_precomp,
                          int32_t* radii, GeomState& g, cudaStream_t s) {
  const float focal_y = prm.H / (2.0f * prm.tanfovy);
  const float focal_x = prm.W / (2.0f * prm.tanfovx);
  const uint32_t gx = (prm.W + TILE_X - 1) / TILE_X, gy = (prm.H + TILE_Y - 1) / TILE_Y;
  const bool use_sh = colors_precomp == nullptr && prm.D > 0 && prm.M > 0;
  size_t smem = use_sh ? (size_t)PRE_THREADS * 3 * prm.M * sizeof(float) + 16 : 0;
  // parameter slots of the cooperative rect walk: 48 bytes per thread, inside the warp's own SH rows when those
  // are large enough (3 M floats per thread >= 48 bytes), else behind the slice
  uint32_t walk_off = 0xffffffffu;
  if (!use_sh || (size_t)3 * prm.M * sizeof(float) < 48) {
    walk_off = (uint32_t)((smem + 15) & ~(size_t)15);
    smem = walk_off + (size_t)PRE_THREADS * 48;
  }
  static bool attr_set[64] = {};
  allow_big_smem(preprocess_fwd_kernel, attr_set);
  if (smem > 96 * 1024) {
    set_error("preprocess: SH row too large (M=%d)", prm.M);
    return 1;
  }
  preprocess_fwd_kernel<<<(prm.P + PRE_THREADS - 1) / PRE_THREADS, PRE_THREADS, smem, s>>>(
      prm.P, prm.D, prm.M, means3D, scales, prm.scale_modifier, rotations, opacities, dc, sh,
      cov3D_precomp, colors_precomp, prm.viewmatrix, prm.projmatrix, prm.campos, prm.W, prm.H,
      prm.tanfovx, prm.tanfovy, focal_x, focal_y, radii, g.rec, g.clamped, g.tiles, g.depth_key,
      g.emit_rec, g.ctl, gx, gy, prm.prefiltered != 0, prm.antialiasing != 0, prm.raw_inputs != 0, walk_off);
  return check_launch("preprocess_fwd", prm.debug != 0, s);
}

int launch_depth_keys(const DgrParams& prm, const GeomState& g, cudaStream_t s) {
  depth_keys_kernel<<<(prm.P + 255) / 256, 256, 0, s>>>(prm.P, g.rec, g.tiles, g.depth_key);
  return check_launch("depth_keys", prm.debug != 0, s);
}

size_t gather_scan_scratch_bytes(size_t P) {
  if (P <= GS_CHAINED_MAX) return align_up(((P + GS_CHUNK - 1) / GS_CHUNK + 2) * sizeof(uint32_t));
  return scan_scratch_bytes(P);
}

// sorted_emit / sorted_cnt / offsets in depth order and the instance total R (to *total_out)
int launch_gather_scan(const DgrParams& prm, const GeomState& g, const uint32_t* sorted_idx, char* scratch,
                       uint32_t* total_out, cudaStream_t s) {
  if ((size_t)prm.P > GS_CHAINED_MAX) {
    // one 16-byte gather per Gaussian brings everything the emit stage needs into depth order; the scan and
    // the emit kernel then stream sequentially
    // (g.offsets holds the scanned sums of the emit kernel's 256-Gaussian blocks here, not per-Gaussian offsets)
    launch_pdl(gather_emit_kernel, dim3((prm.P + 255) / 256), dim3(256), 0, s, (int)prm.P, sorted_idx,
               (const uint4*)g.emit_rec, g.sorted_emit, g.sorted_cnt, g.offsets);
    return launch_block_base_scan(g.offsets, (uint32_t)((prm.P + 255) / 256), total_out, prm.debug != 0, s);
  }
  const uint32_t nchunks = (prm.P + GS_CHUNK - 1) / GS_CHUNK;
  uint32_t* status = reinterpret_cast<uint32_t*>(scratch);
  cudaMemsetAsync(scratch, 0, gather_scan_scratch_bytes(prm.P), s);
  launch_pdl(gather_scan_kernel, dim3(nchunks), dim3(GS_THREADS), 0, s, (int)prm.P, sorted_idx, (const uint4*)g.emit_rec,
             g.sorted_emit, g.sorted_cnt, g.offsets, status, status + nchunks, total_out);
  return check_launch("gather_scan", prm.debug != 0, s);
}

int launch_emit(const DgrParams& prm, const GeomState& g, const int32_t* radii, uint32_t* tile_keys,
                uint32_t* vals, uint32_t capacity, cudaStream_t s) {
  const uint32_t gx = (prm.W + TILE_X - 1) / TILE_X, gy = (prm.H + TILE_Y - 1) / TILE_Y;
  launch_pdl(emit_instances_kernel, dim3((prm.P + 255) / 256), dim3(256), 0, s, (int)prm.P, (const uint4*)g.sorted_emit,
             (const uint32_t*)g.sorted_cnt, (const uint32_t*)g.offsets,
             (size_t)prm.P > GS_CHAINED_MAX ? (const uint32_t*)g.offsets : (const uint32_t*)nullptr,
             (const SplatRecord*)g.rec, (const int*)radii, tile_keys, vals, gx, gy, capacity);
  return check_launch("emit_instances", prm.debug != 0, s);
}

int launch_tile_ranges(const DgrParams& prm, uint32_t R, const uint32_t* R_dev, const uint32_t* tile_keys,
                       uint2* ranges, cudaStream_t s) {
  const uint32_t gx = (prm.W + TILE_X - 1) / TILE_X, gy = (prm.H + TILE_Y - 1) / TILE_Y;
  cudaMemsetAsync(ranges, 0, (size_t)gx * gy * sizeof(uint2), s);
  if (R > 0) launch_pdl(tile_ranges_kernel, dim3((R + 256 * TR_KEYS - 1) / (256 * TR_KEYS)), dim3(256), 0, s, R, R_dev, tile_keys, ranges, gx * gy);
  return check_launch("tile_ranges", prm.debug != 0, s);
}

__global__ void tile_identity_kernel(uint32_t T, uint32_t* __restrict__ order) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < T) order[t] = t;
}

int launch_tile_order(const DgrParams& prm, const ImgState& im, cudaStream_t s) {
  const uint32_t gx = (prm.W + TILE_X - 1) / TILE_X, gy = (prm.H + TILE_Y - 1) / TILE_Y;
  static const bool by_length = !(getenv("DOGS_B200_TILE_ORDER") && getenv("DOGS_B200_TILE_ORDER")[0] == '0');
  if (by_length) launch_pdl(tile_order_kernel, dim3(1), dim3(1024), 0, s, gx * gy, (const uint2*)im.ranges, im.tile_order);
  else tile_identity_kernel<<<(gx * gy + 255) / 256, 256, 0, s>>>(gx * gy, im.tile_order);
  return check_launch("tile_order", prm.debug != 0, s);
}

__global__ void activations_kernel(int P, const float* __restrict__ raw_op, const float* __restrict__ raw_sc,
                                   const float* __restrict__ raw_rot, float* __restrict__ op,
                                   float* __restrict__ sc, float* __restrict__ rot) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= P) return;
  if (raw_op) op[idx] = act_sigmoid(raw_op[idx]);
  if (raw_sc) {
    const float3 s = act_exp3({raw_sc[3 * idx], raw_sc[3 * idx + 1], raw_sc[3 * idx + 2]});
    sc[3 * idx] = s.x; sc[3 * idx + 1] = s.y; sc[3 * idx + 2] = s.z;
  }
  if (raw_rot) reinterpret_cast<float4*>(rot)[idx] = act_normalize(reinterpret_cast<const float4*>(raw_rot)[idx]);
}

int launch_activations(int P, const float* raw_op, const float* raw_sc, const float* raw_rot, float* op,
                       float* sc, float* rot, cudaStream_t s) {
  activations_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, raw_op, raw_sc, raw_rot, op, sc, rot);
  return check_launch("activations", false, s);
}

int launch_filter(const DgrParams& prm, const float* means3D, const float* scales,
                  const float* rotations, const float* cov3D_precomp, int32_t* radii,
                  cudaStream_t s) {
  const float focal_y = prm.H / (2.0f * prm.tanfovy);
  const float focal_x = prm.W / (2.0f * prm.tanfovx);
  const uint32_t gx = (prm.W + TILE_X - 1) / TILE_X, gy = (prm.H + TILE_Y - 1) / TILE_Y;
  filter_kernel<<<(prm.P + 255) / 256, 256, 0, s>>>(
      prm.P, means3D, scales, prm.scale_modifier, rotations, cov3D_precomp, prm.viewmatrix,
      prm.projmatrix, prm.W, prm.H, prm.tanfovx, prm.tanfovy, focal_x, focal_y, radii, gx, gy,
      prm.prefiltered != 0, prm.raw_inputs != 0);
  return check_launch("visible_filter", prm.debug != 0, s);
}

int launch_mark_visible(int P, const float* means3D, const float* view, uint8_t* present,
                        cudaStream_t s) {
  mark_visible_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, means3D, view, present);
  return check_launch("mark_visible", false, s);
}

}  // namespace dogs
