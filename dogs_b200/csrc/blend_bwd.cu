// blend_bwd_kernel — backward of the per-tile alpha blend (replaces PerGaussianRenderCUDA,
// cuda_rasterizer/backward.cu:453-658).
//
// The reference walks each 32-splat bucket over the tile's 256 pixels and needs a (T, colour, depth)
// snapshot per pixel per bucket from the forward pass (5 KB per bucket, ~1.6 GB at 3 M Gaussians).
// Here every pixel replays its own front-to-back blend: T evolves bit-identically to the forward
// pass and the entry count is bounded by n_contrib, so the forward stores no snapshot at all.
//
// Work layout (second version; ncu of the first version — one pixel per thread — showed the
// kernel issue-bound with ~45 % of its instructions in the cross-lane reduction of the 10
// gradient terms, profiles/r1_full_v1_kernels.md):
//   * one CTA of 64 threads per 16x16 tile; a warp owns an 8x16 half tile, each thread 4 pixels
//     (one in each 8x4 block of the half tile).  The four pixels of a thread share the record
//     loads and are summed in registers before anything crosses lanes, so the reduction runs once
//     per (half tile, splat) instead of once per (32 pixels, splat), while the heavy gradient
//     math still runs per compact 8x4 block and is skipped for blocks the splat does not reach;
//   * the colour/depth prefix of the reference (ar[3], ard) is carried as ONE scalar
//     S = sum_c ar_c * dL/dC_c + ard * dL/dD, which removes 3/4 of the per-pair state updates:
//       dL/dalpha = T * (c . dL/dC + invd * dL/dD) + S / (1 - alpha)  - T_final * bg.dL/dC / (1 - alpha);
//   * 10 totals reduced with a transposed butterfly (12 shuffles), stored per warp in shared memory
//     (a warp meets an entry once per batch, so no shared atomic is needed), the two half tiles are
//     summed at the end of the batch and flushed with red.global.add.v4/v2.f32 (3 vector reductions
//     per tile instance).
#include "blend_common.cuh"

#ifndef BB_MIN_CTAS
#define BB_MIN_CTAS 12
#endif

namespace dogs {

constexpr int BB_THREADS = 64;   // 2 warps per tile
constexpr int BB_BATCH = 64;     // records staged per round (one per thread)
constexpr int BB_ACC = 12;       // floats per Gaussian in the global accumulator (3 x float4)

__device__ __forceinline__ void bb_cp_async16(void* smem, const void* gmem) {
  uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void bb_cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void bb_cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}
__device__ __forceinline__ void bb_red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}
__device__ __forceinline__ void bb_red_add_v2(float* p, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}

// Sum 10 per-lane values over the warp.  Afterwards lane l holds the total of value slot(l)
// (see bb_slot); every total lives in two lanes (bit 0 of the lane id).
__device__ __forceinline__ float bb_warp_reduce10(const float v[10], int lane) {
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
__device__ __forceinline__ int bb_slot(int lane) {
  const int b4 = (lane >> 4) & 1, b3 = (lane >> 3) & 1, b2 = (lane >> 2) & 1, b1 = (lane >> 1) & 1;
  if (lane & 1) return -1;
  const int i2 = 2 * b2 + b1;
  if (i2 > 2) return -1;
  if (b3 && i2 > 1) return -1;
  return 5 * b4 + 3 * b3 + i2;
}

// global accumulator slots: 0 mean2D.x, 1 mean2D.y, 2 conic.x, 3 conic.y, 4 conic.w, 5 opacity,
//                           6..8 colour rgb, 9 inverse depth
// shared (per tile) slots 0..5 hold sum q, q*(dx*cx+dy*cy), q*(dy*cz+dx*cy), q*dx^2, q*dx*dy, q*dy^2 instead
// HAS_DEPTH = false: no gradient arrives through the inverse-depth image (the usual photometric loss),
// so the depth terms are compiled out (they would only add exact zeros).
template <bool HAS_DEPTH>
__global__ void __launch_bounds__(BB_THREADS, BB_MIN_CTAS)
blend_bwd_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ tile_order,
                 const uint32_t* __restrict__ point_list, int W,
                 int H, const float* __restrict__ bg_color, const SplatRecord* __restrict__ rec,
                 const float* __restrict__ final_Ts, const uint32_t* __restrict__ n_contrib,
                 const float* __restrict__ pixel_colors, const float* __restrict__ pixel_invdepths,
                 const float* __restrict__ dL_dpixels, const float* __restrict__ dL_invdepths,
                 float* __restrict__ accum, bool ref_quirk, const uint32_t* __restrict__ hits) {
  __shared__ SplatRecord sR[2][BB_BATCH];      // one 48-byte struct per entry: one address for its three loads
  __shared__ uint32_t sId[2][BB_BATCH];
  __shared__ __align__(16) uint32_t sHit[2][16];   // [buffer][2 groups of 32 entries x 8 blocks] from the forward pass
  // Per-warp totals of the batch: a warp meets an entry at most once per round, so it simply STORES its
  // 10 totals (no shared-memory float atomic, which is a compare-and-swap loop on this hardware); sWritten
  // says which entries each warp wrote, the flush adds the two halves.
  __shared__ float sAcc[2][BB_BATCH][11];
  __shared__ uint32_t sWritten[2][2];
  __shared__ uint32_t sMax, sMaxRows01;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // warp w covers tile columns 8w..8w+7 (8 wide, 16 tall); pixel k of lane (lx, ly) = (lane & 7, lane >> 3)
  // is (8w + lx, 4k + ly): the 32 pixels that a warp treats together in step k form a compact 8x4
  // block, so a splat that misses the block costs the warp nothing but the test.
  const uint32_t horizontal_blocks = (W + TILE_X - 1) / TILE_X;
  const uint32_t tile = tile_order[blockIdx.x];     // longest list first (tile_order_kernel)
  const uint32_t tile_x = tile % horizontal_blocks, tile_y = tile / horizontal_blocks;
  const uint32_t qx = tile_x * TILE_X + 8 * warp + (lane & 7);
  const uint32_t qy = tile_y * TILE_Y + (lane >> 3);
  const size_t HW = (size_t)H * W;
  const uint2 range = ranges[tile];
  const float bg0 = bg_color[0], bg1 = bg_color[1], bg2 = bg_color[2];

  // per-pixel state (k = 8x4 block index, top to bottom)
  // (the four pixels of a thread share their x; y = pyf0 + 4k.  S carries the background term as well:
  //  dL/dalpha = cd*T + (S - T_final * bg.dL/dC) / (1 - alpha), and the subtrahend is constant per pixel)
  uint32_t last[4];
  float T[4], S[4], dC0[4], dC1[4], dC2[4], dD[4];
  const float pxf = (float)qx, pyf0 = (float)qy;
  uint32_t lmax = 0, lmax01 = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const uint32_t px = qx, py = qy + 4 * k;
    const bool inside = px < (uint32_t)W && py < (uint32_t)H;
    last[k] = 0; T[k] = 1.0f; S[k] = 0.f; dC0[k] = dC1[k] = dC2[k] = dD[k] = 0.f;
    if (inside) {
      const uint32_t pix_id = W * py + px;
      last[k] = n_contrib[pix_id];
      dC0[k] = dL_dpixels[0 * HW + pix_id];
      dC1[k] = dL_dpixels[1 * HW + pix_id];
      dC2[k] = dL_dpixels[2 * HW + pix_id];
      dD[k] = HAS_DEPTH ? dL_invdepths[pix_id] : 0.f;
      // S = sum_c (prefix_c - pixel_c) * dL/dC_c + (prefix_d - pixel_d) * dL/dD, prefix = 0 at the start
      S[k] = -(pixel_colors[0 * HW + pix_id] * dC0[k] + pixel_colors[1 * HW + pix_id] * dC1[k] +
               pixel_colors[2 * HW + pix_id] * dC2[k] + (HAS_DEPTH ? pixel_invdepths[pix_id] * dD[k] : 0.f));
      // backward.cu:604-621: the background term (-T_final/(1-alpha)) * (bg . dL/dC)
      S[k] -= final_Ts[pix_id] * (bg0 * dC0[k] + bg1 * dC1[k] + bg2 * dC2[k]);
    }
    lmax = max(lmax, last[k]);
    // rows 0-1 of the tile = the first warp of the reference's 16x16 forward block
    if (k == 0 && (lane >> 3) < 2) lmax01 = max(lmax01, last[k]);
  }
  const float ddelx_dx = 0.5 * W;
  const float ddely_dy = 0.5 * H;

  if (tid == 0) { sMax = 0; sMaxRows01 = 0; }
  __syncthreads();
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lmax = max(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
    lmax01 = max(lmax01, __shfl_xor_sync(0xffffffffu, lmax01, o));
  }
  if (lane == 0) {
    atomicMax(&sMax, lmax);
    if (ref_quirk) atomicMax(&sMaxRows01, lmax01);
  }
  __syncthreads();
  uint32_t todo_total = min(range.y - range.x, sMax);  // backward.cu:505-507 with the intended max
  if (ref_quirk) {
    // Reference behaviour (DgrParams::ref_bucket_cull_quirk): 32-entry buckets whose first entry is
    // >= max(n_contrib over rows 0-1) contribute nothing.
    const uint32_t cut = (sMaxRows01 + 31u) & ~31u;
#pragma unroll
    for (int k = 0; k < 4; k++) last[k] = min(last[k], cut);
    todo_total = min(todo_total, cut);
  }
  const int rounds = (todo_total + BB_BATCH - 1) / BB_BATCH;
  const int my_slot = bb_slot(lane);
  // blast[k] = last list position any pixel of the warp's k-th 8x4 block still blends (after the quirk
  // cut): entries at or beyond it are dead for the whole block and are dropped with the geometric cull
  uint32_t blast[4];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    uint32_t v = last[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    blast[k] = v;
  }

  const uint32_t* const hit_in = hits + ((size_t)(range.x / 32) + tile) * 8;
  auto fetch = [&](int round, int buf) {
    uint32_t progress = round * BB_BATCH + tid;
    if (progress < todo_total) {
      uint32_t id = point_list[range.x + progress];
      sId[buf][tid] = id;
      const SplatRecord* r = rec + id;
      bb_cp_async16(&sR[buf][tid].a, &r->a);
      bb_cp_async16(&sR[buf][tid].b, &r->b);
      bb_cp_async16(&sR[buf][tid].c, &r->c);
    }
    // the forward's reachability words of these 64 entries: 2 groups x 8 blocks = 64 contiguous bytes
    if (tid < 4) bb_cp_async16(&sHit[buf][4 * tid], hit_in + (size_t)round * 16 + 4 * tid);
    bb_cp_async_commit();
  };
  if (rounds > 0) fetch(0, 0);

  for (int i = 0; i < rounds; i++) {
    const int buf = i & 1;
    if (i + 1 < rounds) fetch(i + 1, buf ^ 1); else bb_cp_async_commit();
    bb_cp_async_wait<1>();
    __syncthreads();
    const int nbatch = min(BB_BATCH, (int)todo_total - i * BB_BATCH);
    const uint32_t base = i * BB_BATCH;

    // Sub-tile culling: which entries of the batch can reach each of the four 8x4 blocks of this warp's half
    // tile.  The forward pass made that test for the same blocks (its warp 2k + w owns block k of half tile w)
    // and left one word per (group of 32 entries, block) in BinState::hits, so the backward only drops the
    // entries no pixel of the block blends any more (list position >= blast[k]).  (First version: the test
    // was repeated here, one entry per lane and four blocks per entry: 9 % of the kernel's instructions.)
#pragma unroll
    for (int w = 0; w < 2; w++) {
    // lane l keeps the 4-bit block mask of entry 32w + l; m = entries that reach any block
    uint32_t bm = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const bool alive = base + 32 * w + lane < blast[k];
      bm |= (uint32_t)(alive && ((sHit[buf][8 * w + 2 * k + warp] >> lane) & 1u)) << k;
    }
    // (a ballot rather than the OR of the words: the compiler then knows m is warp-uniform and keeps the
    //  loop control on the uniform datapath — 10 % fewer instructions in this kernel)
    uint32_t m = __ballot_sync(0xffffffffu, bm != 0);
    uint32_t wrote = 0;
    while (m) {
      const int bit = __ffs(m) - 1;
      m &= m - 1;
      const int j = 32 * w + bit;
      const uint32_t e = base + j;
      const uint32_t mk = __shfl_sync(0xffffffffu, bm, bit);   // warp-uniform block mask of entry j
      const SplatRecord& ent = sR[buf][j];
      const float4 con_o = ent.a;
      const float4 b = ent.b;
      const float4 cc = ent.c;
      const float invd = b.z;
      const float dx = b.x - pxf;
      float v[10];
#pragma unroll
      for (int q = 0; q < 10; q++) v[q] = 0.f;
      bool any_c = false;
#pragma unroll
      for (int k = 0; k < 4; k++) {
        if (!(mk & (1u << k))) continue;                    // warp-uniform: the splat cannot reach this 8x4 block
        // power, expf and the alpha threshold are the forward's own expressions, so the contribution
        // decisions and T are those of the forward pass bit for bit.  (A fast ex2.approx exp was
        // measured again with the FP64 chain rule in place: 1.069 -> 0.991 ms at C3 and the errors against
        // the FP64 truth stay at 1e-5 — except where the approximate alpha falls on the other side of the
        // 1/255 threshold than the forward's: the pixel's S then disagrees with the forward's colour and a
        // handful of elements (3 at C2) leave the 1e-4 bar.  Re-evaluating expf under a warp vote when a
        // lane is within the approximation error of the threshold restores the errors exactly, but the
        // vote + compare per (block, entry) and 3 more registers cost more than the exp saves: 1.113 ms.)
        const float dy = b.y - (pyf0 + 4.f * k);
        const float power = -0.5f * (con_o.x * dx * dx + con_o.z * dy * dy) - con_o.y * dx * dy;
        const float G = exp(power);
        const float al = min(0.99f, con_o.w * G);
        const bool c = (e < last[k]) && !(power > 0.0f) && !(al < 1.0f / 255.0f);   // backward.cu:592-600
        if (!__any_sync(0xffffffffu, c)) continue;
        any_c = true;
        const float alpha = c ? al : 0.f;              // masked lanes: weight 0, T and S unchanged
        const float Gk = c ? G : 0.f;                  // (exp(power) may be inf where power > 0)
        // c . dL/dC + invd * dL/dD
        const float cd = HAS_DEPTH ? b.w * dC0[k] + cc.x * dC1[k] + cc.y * dC2[k] + invd * dD[k]
                                   : b.w * dC0[k] + cc.x * dC1[k] + cc.y * dC2[k];
        const float Tk = T[k];
        const float weight = alpha * Tk;
        const float om = 1.0f - alpha;
        // om is in [0.01, 1]: one MUFU.RCP (<= 1 ulp; gradients carry a 1e-4 bar).  __fdividef(1, om) wraps the
        // same instruction in a range fix-up for huge divisors (7 instructions, 7 % of this kernel).
        float inv;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(om));
        S[k] += weight * cd;
        const float dL_dalpha = c ? (cd * Tk + inv * S[k]) : 0.f;
        T[k] = Tk * om;
        v[6] += weight * dC0[k];
        v[7] += weight * dC1[k];
        v[8] += weight * dC2[k];
        if (HAS_DEPTH) v[9] += weight * dD[k];
        // Only pixel-dependent moments of q = G * dL/dalpha are accumulated here; the per-splat
        // factors (opacity, conic, 0.5*W, 0.5*H) of backward.cu:626-641 are applied once per
        // (tile, splat) when the batch is flushed.
        // (The conic is applied to the mean terms per pixel: dx*cx + dy*cy cancels strongly for
        // elongated splats, and summing the two moments separately loses that cancellation.)
        const float q = Gk * dL_dalpha;
        const float qx = q * dx, qy = q * dy;
        v[0] += q;
        v[1] += q * (dx * con_o.x + dy * con_o.y);
        v[2] += q * (dy * con_o.z + dx * con_o.y);
        v[3] += qx * dx;
        v[4] += qx * dy;
        v[5] += qy * dy;
      }
      if (!any_c) continue;
      const float tot = bb_warp_reduce10(v, lane);
      if (my_slot >= 0) sAcc[warp][j][my_slot] = tot;
      wrote |= 1u << bit;
    }
    if (lane == 0) sWritten[warp][w] = wrote;
    }
    __syncthreads();
    // flush the batch: thread t owns splat t
    if (tid < nbatch) {
      float a[10];
      const bool w0 = (sWritten[0][tid >> 5] >> (tid & 31)) & 1u, w1 = (sWritten[1][tid >> 5] >> (tid & 31)) & 1u;
#pragma unroll
      for (int q = 0; q < 10; q++) a[q] = (w0 ? sAcc[0][tid][q] : 0.f) + (w1 ? sAcc[1][tid][q] : 0.f);
      if (w0 | w1) {
        // moments -> gradients (backward.cu:626-641):  dL/dG * G = o * q
        const float4 co = sR[buf][tid].a;
        const float o = co.w;
        const float g_mx = -o * ddelx_dx * a[1];
        const float g_my = -o * ddely_dy * a[2];
        const float h = -0.5f * o;
        float* dst = accum + (size_t)sId[buf][tid] * BB_ACC;
        bb_red_add_v4(dst, g_mx, g_my, h * a[3], h * a[4]);
        bb_red_add_v4(dst + 4, h * a[5], a[0], a[6], a[7]);
        bb_red_add_v2(dst + 8, a[8], a[9]);
      }
    }
    __syncthreads();
  }
  bb_cp_async_wait<0>();
}

size_t bwd_accum_bytes(size_t P) { return align_up(P * BB_ACC * sizeof(float)); }

int launch_blend_bwd(const DgrParams& prm, const GeomState& g, const BinState& b, const ImgState& im,
                     const float* dL_dcolor, const float* dL_dinvdepth, float* accum,
                     cudaStream_t s) {
  cudaMemsetAsync(accum, 0, (size_t)prm.P * BB_ACC * sizeof(float), s);
  const uint32_t grid = ((prm.W + TILE_X - 1) / TILE_X) * ((prm.H + TILE_Y - 1) / TILE_Y);
  if (dL_dinvdepth)
    blend_bwd_kernel<true><<<grid, BB_THREADS, 0, s>>>(im.ranges, im.tile_order, b.point_list, prm.W, prm.H, prm.bg, g.rec,
                                               im.final_T, im.n_contrib, im.pix_color,
                                               im.pix_invdepth, dL_dcolor, dL_dinvdepth, accum,
                                               prm.ref_bucket_cull_quirk != 0, b.hits);
  else
    blend_bwd_kernel<false><<<grid, BB_THREADS, 0, s>>>(im.ranges, im.tile_order, b.point_list, prm.W, prm.H, prm.bg, g.rec,
                                               im.final_T, im.n_contrib, im.pix_color,
                                               im.pix_invdepth, dL_dcolor, dL_dinvdepth, accum,
                                               prm.ref_bucket_cull_quirk != 0, b.hits);
  return check_launch("blend_bwd", prm.debug != 0, s);
}

}  // namespace dogs
