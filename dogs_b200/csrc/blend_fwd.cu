// blend_fwd_kernel — per-tile front-to-back alpha blend (replaces renderCUDA,
// cuda_rasterizer/forward.cu:349-501).  The per-pixel arithmetic (power, expf, alpha clamp, the
// 1/255 and 1e-4 thresholds, colour / inverse-depth accumulation) is the reference's, expression
// for expression, so n_contrib, final_T and the image are bit-identical.
//
// What differs:
//   * one CTA (256 threads) per 16x16 tile, but each warp owns a compact 8x4 pixel block and, per
//     staged batch, first culls the list against its own block with one entry per lane
//     (block_reachable, blend_common.cuh); the per-pixel loop only visits the surviving entries.
//     ncu of the first version showed the kernel instruction-issue bound (91 % issue-active) with
//     every pixel testing every list entry;
//   * batches of 256 records are gathered global->shared with cp.async into a double buffer, so
//     the gather of batch i+1 overlaps the blend of batch i (the reference loads colour and depth
//     from global memory inside the inner loop, forward.cu:471,473);
//   * no per-bucket (T, colour, depth) snapshots are written: the backward replays the blend from
//     final_T / n_contrib (blend_bwd.cu).
#include "blend_common.cuh"

#ifndef BF_MIN_CTAS
#define BF_MIN_CTAS 6
#endif

namespace dogs {

// COUNT = true additionally counts, per Gaussian, the pixels it was blended into (LightGaussian
// importance statistics: renderCUDA_count of the reference's older rasterizer,
// submodules/old_diff-gaussian-rasterization/cuda_rasterizer/forward.cu:400-513, whose
// `gaussian_count[id]++` is a non-atomic race; here the lanes of a warp are tallied with a ballot,
// tiles add up in shared memory and one integer atomic per (tile, Gaussian) reaches global memory).
template <bool COUNT>
__global__ void __launch_bounds__(TILE_PIX, BF_MIN_CTAS)
blend_fwd_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ tile_order,
                 const uint32_t* __restrict__ point_list, int W,
                 int H, const SplatRecord* __restrict__ rec, const float* __restrict__ bg_color,
                 float* __restrict__ final_T, uint32_t* __restrict__ n_contrib,
                 float* __restrict__ out_color, float* __restrict__ invdepth,
                 float* __restrict__ pix_color, float* __restrict__ pix_invdepth,
                 int* __restrict__ gaussian_count, uint32_t* __restrict__ hits) {
  // staged records, one 48-byte struct per entry: the three 16-byte loads of an entry share one address computation
  __shared__ SplatRecord sR[2][TILE_PIX];
  __shared__ uint32_t sId[COUNT ? 2 : 1][COUNT ? TILE_PIX : 1];
  __shared__ int sCnt[COUNT ? TILE_PIX : 1];
  __shared__ uint32_t sHit[2][TILE_PIX / 32][8];   // [buffer][group of 32 entries][8x4 block = warp]
  if (COUNT) sCnt[threadIdx.x] = 0;
  pdl_wait();
  pdl_trigger();
  const uint32_t tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const uint32_t horizontal_blocks = (W + TILE_X - 1) / TILE_X;
  // CTAs are launched longest list first (tile_order_kernel)
  const uint32_t tile = tile_order[blockIdx.x];
  const uint32_t tile_x = tile % horizontal_blocks, tile_y = tile / horizontal_blocks;
  // warp -> 8x4 block (2 columns x 4 rows of blocks per tile); lane -> pixel inside the block
  const uint32_t bx0 = tile_x * TILE_X + (warp & 1) * 8, by0 = tile_y * TILE_Y + (warp >> 1) * 4;
  const uint2 pix = {bx0 + (lane & 7), by0 + (lane >> 3)};
  const uint32_t pix_id = W * pix.y + pix.x;
  const float2 pixf = {(float)pix.x, (float)pix.y};
  const bool inside = pix.x < (uint32_t)W && pix.y < (uint32_t)H;
  bool done = !inside;
  const float rx0 = (float)bx0, ry0 = (float)by0, rx1 = (float)(bx0 + 7), ry1 = (float)(by0 + 3);

  const uint2 range = ranges[tile];
  const int rounds = ((range.y - range.x + TILE_PIX - 1) / TILE_PIX);
  int toDo = range.y - range.x;

  float T = 1.0f;
  uint32_t last_contributor = 0;
  float C[3] = {0.f, 0.f, 0.f};
  float expected_invdepth = 0.0f;

  auto fetch = [&](int round, int buf) {
    uint32_t progress = range.x + round * TILE_PIX + tid;
    if (progress < range.y) {
      const uint32_t id = point_list[progress];
      if (COUNT) sId[buf][tid] = id;
      const SplatRecord* r = rec + id;
      cp_async16(&sR[buf][tid].a, &r->a);
      cp_async16(&sR[buf][tid].b, &r->b);
      cp_async16(&sR[buf][tid].c, &r->c);
    }
    cp_async_commit();
  };
  if (rounds > 0) fetch(0, 0);

  // The reachability words of a batch go out to the backward pass (BinState::hits) one barrier later, as one
  // contiguous store of up to 256 bytes by the first 64 threads.  (Nothing is carried across the blend loop for
  // it: the kernel runs on 32 registers.)
  auto flush_hits = [&](int round, int left) {       // left = list entries from the start of `round` on
    const int groups = (min(TILE_PIX, left) + 31) / 32;
    if ((int)tid < groups * 8)
      hits[((size_t)(range.x / 32) + tile + (size_t)round * (TILE_PIX / 32)) * 8 + tid] = sHit[round & 1][tid >> 3][tid & 7];
  };

  int i = 0;
  for (; i < rounds; i++, toDo -= TILE_PIX) {
    const int buf = i & 1;
    // forward.cu:415-417 (the barrier also frees buffer buf^1, last read in round i-1)
    int num_done = __syncthreads_count(done);
    if (i > 0) flush_hits(i - 1, toDo + TILE_PIX);
    if (num_done == TILE_PIX) break;
    if (i + 1 < rounds) fetch(i + 1, buf ^ 1); else cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const int nbatch = min(TILE_PIX, toDo);
    const uint32_t base = i * TILE_PIX;

    // which entries of the batch can reach this warp's 8x4 block (one entry per lane)
    uint32_t hit[TILE_PIX / 32];
#pragma unroll
    for (int r = 0; r < TILE_PIX / 32; r++) {
      const int j = r * 32 + lane;
      bool h = false;
      if (j < nbatch) {
        const SplatRecord& e = sR[buf][j];
        const float4 b = e.b;
        h = block_reachable(e.a, b.x, b.y, rx0, ry0, rx1, ry1, e.c.w);
      }
      hit[r] = __ballot_sync(0xffffffffu, h);
      if (lane == r) sHit[buf][r][warp] = hit[r];
    }

#pragma unroll
    for (int r = 0; r < TILE_PIX / 32; r++) {
      uint32_t m = hit[r];
      // every pixel of the block saturated: nothing left to do for it.  (Checked once per 32 entries: a vote
      // per entry was 6 % of the kernel's instructions; finished lanes skip the work below anyway.)
      if (__all_sync(0xffffffffu, done)) break;
      while (m) {
        const int j = r * 32 + __ffs(m) - 1;
        m &= m - 1;
        bool blended = false;
        {
          // One predicate instead of four nested branches: power, exp and alpha are evaluated by every lane (a lane
          // that is done, or whose power is positive, computes a value nobody uses), the state changes only under ok.
          const SplatRecord& e = sR[buf][j];
          const float4 con_o = e.a;
          const float4 b = e.b;
          float2 xy = {b.x, b.y};
          float2 d = {xy.x - pixf.x, xy.y - pixf.y};
          float power = -0.5f * (con_o.x * d.x * d.x + con_o.z * d.y * d.y) - con_o.y * d.x * d.y;
          float alpha = min(0.99f, con_o.w * exp(power));
          const bool ok = !done && !(power > 0.0f) && !(alpha < 1.0f / 255.0f);
          if (ok) {
            float test_T = T * (1 - alpha);
            if (test_T < 0.0001f) {
              done = true;
            } else {
              const float4 c = e.c;
              C[0] += b.w * alpha * T;
              C[1] += c.x * alpha * T;
              C[2] += c.y * alpha * T;
              expected_invdepth += b.z * alpha * T;
              T = test_T;
              last_contributor = base + j + 1;   // forward.cu:444,479: position in the tile's list
              blended = true;
            }
          }
        }
        if (COUNT) {
          const uint32_t n = __popc(__ballot_sync(0xffffffffu, blended));
          if (lane == 0 && n) atomicAdd(&sCnt[j], (int)n);
        }
      }
    }
    if (COUNT) {   // flush this batch's tallies: thread t owns entry t
      __syncthreads();
      if ((int)tid < nbatch && sCnt[tid]) atomicAdd(&gaussian_count[sId[buf][tid]], sCnt[tid]);
      sCnt[tid] = 0;
    }
  }
  cp_async_wait<0>();
  if (i == rounds && rounds > 0) {      // the loop ran to its end: the last batch is still in shared memory
    __syncthreads();
    flush_hits(rounds - 1, toDo + TILE_PIX);
  }

  if (inside) {
    final_T[pix_id] = T;
    n_contrib[pix_id] = last_contributor;
    const size_t HW = (size_t)H * W;
#pragma unroll
    for (int ch = 0; ch < 3; ch++) {
      const float v = C[ch] + T * bg_color[ch];
      out_color[ch * HW + pix_id] = v;
      pix_color[ch * HW + pix_id] = v;  // rasterizer_impl.cu:495-496 (copy kept for backward)
    }
    invdepth[pix_id] = expected_invdepth;
    pix_invdepth[pix_id] = expected_invdepth;
  }
}

// important_score[id] = sum over the blended pixels of conic_opacity.w (old forward.cu:486), i.e.
// count * opacity; written for every Gaussian (0 where count == 0).
__global__ void importance_kernel(int P, const int* __restrict__ count, const SplatRecord* __restrict__ rec,
                                  float* __restrict__ score) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  const int c = count[i];
  score[i] = c ? (float)c * rec[i].a.w : 0.f;
}

int launch_blend_fwd(const DgrParams& prm, const GeomState& g, const BinState& b, const ImgState& im,
                     float* out_color, float* out_invdepth, cudaStream_t s) {
  const uint32_t grid = ((prm.W + TILE_X - 1) / TILE_X) * ((prm.H + TILE_Y - 1) / TILE_Y);
  if (prm.gaussians_count) {
    cudaMemsetAsync(prm.gaussians_count, 0, (size_t)prm.P * sizeof(int32_t), s);
    blend_fwd_kernel<true><<<grid, TILE_PIX, 0, s>>>(im.ranges, im.tile_order, b.point_list, prm.W, prm.H, g.rec, prm.bg,
                                                     im.final_T, im.n_contrib, out_color, out_invdepth,
                                                     im.pix_color, im.pix_invdepth, prm.gaussians_count, b.hits);
    if (prm.important_score)
      importance_kernel<<<(prm.P + 255) / 256, 256, 0, s>>>(prm.P, prm.gaussians_count, g.rec, prm.important_score);
    return check_launch("blend_fwd(count)", prm.debug != 0, s, prm.important_score ? 2 : 1);
  }
  launch_pdl(blend_fwd_kernel<false>, dim3(grid), dim3(TILE_PIX), 0, s, (const uint2*)im.ranges, (const uint32_t*)im.tile_order,
             (const uint32_t*)b.point_list, (int)prm.W, (int)prm.H, (const SplatRecord*)g.rec, prm.bg, im.final_T, im.n_contrib,
             out_color, out_invdepth, im.pix_color, im.pix_invdepth, (int*)nullptr, b.hits);
  return check_launch("blend_fwd", prm.debug != 0, s);
}

}  // namespace dogs
