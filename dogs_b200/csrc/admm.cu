// ADMM Gaussian-consensus kernels.  They replace the torch tensor arithmetic of
//   master_gaussian_trainer.py:538-555 (consensus average), :396-456 (primal/dual residuals),
//   slave_gaussian_trainer.py:99-121 (dual update) and :161-202 (penalty terms)
// for the scene blocks of one GPU.  Only Gaussians owned by more than one block travel: one launch packs
// them as 59-float rows straight into the per-co-owner exchange buffers, and after the exchange one
// launch averages (in the reference's block order, so z is bit-identical to the master's), applies
// the dual update and accumulates the residual sums: the 59 floats of a Gaussian are read and written
// once per consensus round.
//
// A Gaussian = 59 floats in 6 tensors: xyz 3, f_dc 3, f_rest 45, scaling 3, quaternion 4, opacity 1
// (raw, pre-activation values are the consensus variables).
//
// All kernels are HBM-bound element-wise passes: every tensor is walked as a flat array with
// consecutive threads on consecutive elements (16-byte accesses in the per-iteration penalty
// kernel); per-group sums are kept in FP64.
#include "common.cuh"

namespace dogs {

constexpr int ADMM_ROW = 59;
constexpr int ADMM_THREADS = 256;

struct PenaltyArgs {
  const float* x[6];
  const float* u[6];
  const float* z[6];
  float* grad[6];
  unsigned long long chunk_end[6];   // exclusive prefix ends, in float4 chunks
  unsigned long long len[6];         // elements per tensor
  float coef[6];
};

__device__ __forceinline__ double block_sum_to(double v, double* smem_slot) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0 && v != 0.0) atomicAdd(smem_slot, v);
  return v;
}

// One launch for all six tensors: r = x + u - z;  sums[g] += sum r^2;  grad += coef[g] * r.
__global__ void __launch_bounds__(ADMM_THREADS)
admm_penalty_kernel(const __grid_constant__ PenaltyArgs a, double* __restrict__ sums) {
  __shared__ double acc[6];
  if (threadIdx.x < 6) acc[threadIdx.x] = 0.0;
  __syncthreads();
  const unsigned long long total = a.chunk_end[5];
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  double part = 0.0;      // running sum of the group this thread is currently in
  int part_g = 0;
  for (unsigned long long c = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; c < total; c += stride) {
    int g = 0;
#pragma unroll
    for (int q = 0; q < 5; q++) g += (c >= a.chunk_end[q]) ? 1 : 0;
    const unsigned long long lc = c - (g ? a.chunk_end[g - 1] : 0ull);
    const unsigned long long e0 = lc * 4;
    const float coef = a.coef[g];
    const float* __restrict__ xp = a.x[g];
    const float* __restrict__ up = a.u[g];
    const float* __restrict__ zp = a.z[g];
    float* __restrict__ gp = a.grad[g];
    float s = 0.f;
    if (e0 + 4 <= a.len[g]) {
      const float4 xv = *reinterpret_cast<const float4*>(xp + e0);
      const float4 uv = up ? *reinterpret_cast<const float4*>(up + e0) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 zv = *reinterpret_cast<const float4*>(zp + e0);
      const float r0 = (xv.x + uv.x) - zv.x, r1 = (xv.y + uv.y) - zv.y;   // mse_loss(x + u, z), slave:169-187
      const float r2 = (xv.z + uv.z) - zv.z, r3 = (xv.w + uv.w) - zv.w;
      s = r0 * r0 + r1 * r1 + r2 * r2 + r3 * r3;
      if (gp) {
        float4 gv = *reinterpret_cast<float4*>(gp + e0);
        gv.x += coef * r0; gv.y += coef * r1; gv.z += coef * r2; gv.w += coef * r3;   // rho/numel * r
        *reinterpret_cast<float4*>(gp + e0) = gv;
      }
    } else {
      for (unsigned long long e = e0; e < a.len[g]; e++) {
        const float r = (xp[e] + (up ? up[e] : 0.f)) - zp[e];
        s += r * r;
        if (gp) gp[e] += coef * r;
      }
    }
    if (g != part_g) {      // rare: the grid-stride walk crossed into the next tensor
      if (part != 0.0) atomicAdd(&acc[part_g], part);
      part = 0.0;
      part_g = g;
    }
    part += (double)s;
  }
  if (part != 0.0) atomicAdd(&acc[part_g], part);
  __syncthreads();
  if (threadIdx.x < 6 && acc[threadIdx.x] != 0.0) atomicAdd(&sums[threadIdx.x], acc[threadIdx.x]);
}

// ---- consensus round: two launches per block --------------------------------------------------------
// Work layout shared by both kernels: one warp per Gaussian row, lane l owns elements l and l + 32 of
// the 59-float row, so the consensus group of a lane (xyz / f_dc / f_rest / scaling / quaternion /
// opacity) is FIXED: no per-element division, the per-group residual sums stay in registers, and
// every exchange-buffer row is read / written as 236 contiguous bytes.
struct RowLane {
  int g[2], off[2], sz[2];   // tensor, offset in the tensor's row, row length; g = -1: lane has no 2nd element
};
__device__ __forceinline__ RowLane row_lane(int lane) {
  RowLane r;
#pragma unroll
  for (int h = 0; h < 2; h++) {
    const int e = lane + 32 * h;
    int g, off, sz;
    if (e < 3) { g = 0; off = e; sz = 3; }
    else if (e < 6) { g = 1; off = e - 3; sz = 3; }
    else if (e < 51) { g = 2; off = e - 6; sz = 45; }
    else if (e < 54) { g = 3; off = e - 51; sz = 3; }
    else if (e < 58) { g = 4; off = e - 54; sz = 4; }
    else if (e < 59) { g = 5; off = 0; sz = 1; }
    else { g = -1; off = 0; sz = 0; }
    r.g[h] = g; r.off[h] = off; r.sz[h] = sz;
  }
  return r;
}

constexpr int ADMM_MAX_DESTS = 16;
struct PackArgs {
  const float* x[6];
  float* dst[ADMM_MAX_DESTS];   // scatter mode: the buffers rows can go to (own send buffer, peers' receive buffers)
};

// out[r*59 + e] = x[src_local[r]][e] for the R rows this block sends to its co-owners (rows of all peers
// concatenated; a Gaussian shared with three other blocks appears three times).  Replaces the six
// per-tensor gathers + one torch.index_select per peer of the first version.
// dst_sel != null: row r goes to a.dst[dst_sel[r]] at row dst_row[r] — the receive buffer of the co-owner's GPU,
// mapped into this process (peer memory over NVLink): the pack kernel IS the exchange, there is no send buffer
// to copy from and no collective call for these rows.
__global__ void __launch_bounds__(ADMM_THREADS)
admm_pack_rows_kernel(const __grid_constant__ PackArgs a, const int32_t* __restrict__ src_local,
                      const uint8_t* __restrict__ dst_sel, const int32_t* __restrict__ dst_row,
                      int64_t R, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const RowLane rl = row_lane(lane);
  const float* p0 = a.x[rl.g[0]];
  const float* p1 = rl.g[1] >= 0 ? a.x[rl.g[1]] : nullptr;
  const int64_t nwarps = (int64_t)gridDim.x * (ADMM_THREADS / 32);
  for (int64_t r = (int64_t)blockIdx.x * (ADMM_THREADS / 32) + (threadIdx.x >> 5); r < R; r += nwarps) {
    const int64_t li = src_local[r];
    const float v0 = __ldg(p0 + li * rl.sz[0] + rl.off[0]);
    float v1 = 0.f;
    if (p1) v1 = __ldg(p1 + li * rl.sz[1] + rl.off[1]);
    float* o = dst_sel ? a.dst[dst_sel[r]] + (int64_t)dst_row[r] * ADMM_ROW : out + r * ADMM_ROW;
    o[lane] = v0;
    if (p1) o[32 + lane] = v1;
  }
}

constexpr int ADMM_MAX_SOURCES = 16;
struct ConsensusArgs {
  const float* x[6];
  float* u[6];
  float* z[6];
  float* zmu[6];                        // nullable (all or none)
  const float* src[ADMM_MAX_SOURCES];   // exchange buffers [rows, 59] the addends come from
};

// Over ALL n Gaussians of the block:
//   shared (shared_row[i] >= 0): z = (a_0 + a_1 + ...) / count with the addends in ASCENDING BLOCK ORDER
//       (plus_gaussians over k = 0..K-1, then average_gaussians: gaussian_splat_model.py:325-340 called
//       from master_gaussian_trainer.py:549-555), each addend either this block's own x or a row of a
//       co-owner's exchange buffer; true IEEE division by the count;
//       u += relax * (x - z) rounded like torch's three separate ops (slave_gaussian_trainer.py:99-121);
//   single owner: z = x, u unchanged (the reference adds relax * 0).
//   zmu = z - u for the per-iteration penalty.  Residual sums in FP64: partial[blk][g] = sum (z - x)^2
//   (primal, master:396-432), partial[blk][6+g] = sum (z_prev - z)^2 / count (dual, master:435-456: every
//   global Gaussian once); the last CTA to finish folds the partials in a fixed order into residuals[12].
// MAXA > 0: at most MAXA addends per row, all table entries and all addend rows are fetched before anything is
// summed (independent loads in flight instead of an entry -> row -> entry -> row chain); MAXA = 0: any number.
template <int MAXA>
__global__ void __launch_bounds__(ADMM_THREADS)
admm_consensus_kernel(const __grid_constant__ ConsensusArgs a, const int32_t* __restrict__ shared_row,
                      const uint32_t* __restrict__ addends, const uint8_t* __restrict__ count, int max_addends,
                      int64_t n, float relax, double* __restrict__ partial, unsigned int* __restrict__ ticket,
                      double* __restrict__ residuals) {
  __shared__ double acc[12];
  __shared__ double inv_cnt[ADMM_THREADS];     // 1 / count without a double division per element
  __shared__ bool last;
  if (threadIdx.x < 12) acc[threadIdx.x] = 0.0;
  inv_cnt[threadIdx.x] = 1.0 / (double)max((int)threadIdx.x, 1);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const RowLane rl = row_lane(lane);
  const bool has_zmu = a.zmu[0] != nullptr;
  double primal[2] = {0.0, 0.0}, dual[2] = {0.0, 0.0};
  const int64_t nwarps = (int64_t)gridDim.x * (ADMM_THREADS / 32);
  for (int64_t i = (int64_t)blockIdx.x * (ADMM_THREADS / 32) + (threadIdx.x >> 5); i < n; i += nwarps) {
    const int32_t sr = shared_row[i];
    int cnt_i = 1;
    uint32_t ent[MAXA > 0 ? MAXA : 1];
    if (sr >= 0) {
      cnt_i = count[sr];
      if (MAXA > 0) {
#pragma unroll
        for (int j = 0; j < MAXA; j++) ent[j] = j < max_addends ? addends[(int64_t)sr * max_addends + j] : 0xffffffffu;
      }
    }
    const float cnt = (float)cnt_i;
    const double w = inv_cnt[cnt_i];
    float xv[2], z_old[2], u_val[2], add[2][MAXA > 0 ? MAXA : 1];
#pragma unroll
    for (int h = 0; h < 2; h++) {
      if (rl.g[h] < 0) continue;
      const int64_t e = i * rl.sz[h] + rl.off[h];
      xv[h] = a.x[rl.g[h]][e];
      z_old[h] = a.z[rl.g[h]][e];
      u_val[h] = a.u[rl.g[h]][e];
      if (MAXA > 0 && sr >= 0) {
#pragma unroll
        for (int j = 0; j < MAXA; j++) {
          const uint32_t which = ent[j] >> 28;
          add[h][j] = (ent[j] == 0xffffffffu || which == 15u)
                          ? 0.f
                          : __ldg(a.src[which] + (int64_t)(ent[j] & 0x0fffffffu) * ADMM_ROW + lane + 32 * h);
        }
      }
    }
#pragma unroll
    for (int h = 0; h < 2; h++) {
      if (rl.g[h] < 0) continue;
      const int g = rl.g[h];
      const int64_t e = i * rl.sz[h] + rl.off[h];
      float u_new = u_val[h];
      float z_new = xv[h];
      if (sr >= 0) {
        float s = 0.f;
        if (MAXA > 0) {
#pragma unroll
          for (int j = 0; j < MAXA; j++) {
            if (ent[j] == 0xffffffffu) break;
            const float v = (ent[j] >> 28) == 15u ? xv[h] : add[h][j];
            s = j == 0 ? v : __fadd_rn(s, v);
          }
        } else {
          for (int j = 0; j < max_addends; j++) {
            const uint32_t en = addends[(int64_t)sr * max_addends + j];
            if (en == 0xffffffffu) break;
            const uint32_t which = en >> 28;
            const float v = which == 15u ? xv[h]
                                         : __ldg(a.src[which] + (int64_t)(en & 0x0fffffffu) * ADMM_ROW + lane + 32 * h);
            s = j == 0 ? v : __fadd_rn(s, v);
          }
        }
        z_new = __fdiv_rn(s, cnt);
        u_new = __fadd_rn(u_new, __fmul_rn(relax, __fsub_rn(xv[h], z_new)));
        a.u[g][e] = u_new;
      }
      a.z[g][e] = z_new;
      if (has_zmu) a.zmu[g][e] = z_new - u_new;
      const double dp = (double)z_new - (double)xv[h], dd = (double)z_old[h] - (double)z_new;
      primal[h] += dp * dp;
      dual[h] += dd * dd * w;
    }
  }
#pragma unroll
  for (int h = 0; h < 2; h++) {
    if (rl.g[h] < 0) continue;
    if (primal[h] != 0.0) atomicAdd(&acc[rl.g[h]], primal[h]);
    if (dual[h] != 0.0) atomicAdd(&acc[6 + rl.g[h]], dual[h]);
  }
  __syncthreads();
  if (threadIdx.x < 12) partial[(size_t)blockIdx.x * 12 + threadIdx.x] = acc[threadIdx.x];
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (last && threadIdx.x < 12) {
    double t = 0.0;
    for (unsigned b = 0; b < gridDim.x; b++) t += __ldcg(&partial[(size_t)b * 12 + threadIdx.x]);
    residuals[threadIdx.x] += t;
    if (threadIdx.x == 0) *ticket = 0u;
  }
}

static const int kSize[6] = {3, 3, 45, 3, 4, 1};

static int grid_for(unsigned long long work_items) {
  unsigned long long g = (work_items + ADMM_THREADS - 1) / ADMM_THREADS;
  const unsigned long long cap = 148ull * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace dogs

using namespace dogs;

static int consensus_grid(int64_t rows) {
  const int64_t per_cta = ADMM_THREADS / 32;
  int64_t g = (rows + per_cta - 1) / per_cta;
  const int64_t cap = 148 * 8;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

extern "C" size_t admm_consensus_scratch_bytes(void) { return (size_t)(148 * 8) * 12 * sizeof(double) + 256; }

extern "C" int admm_pack_rows(const AdmmConstTensors* x, const int32_t* src_local, int64_t R, float* out,
                              void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (R <= 0) return 0;
  if (!x || !src_local || !out) { set_error("admm_pack_rows: null pointer"); return 1; }
  PackArgs a;
  for (int g = 0; g < 6; g++) {
    a.x[g] = x->g[g];
    if (!a.x[g]) { set_error("admm_pack_rows: null tensor"); return 1; }
  }
  for (int k = 0; k < ADMM_MAX_DESTS; k++) a.dst[k] = nullptr;
  admm_pack_rows_kernel<<<consensus_grid(R), ADMM_THREADS, 0, s>>>(a, src_local, nullptr, nullptr, R, out);
  return check_launch("admm_pack_rows", false, s);
}

extern "C" int admm_pack_rows_to(const AdmmConstTensors* x, const int32_t* src_local, const uint8_t* dst_sel,
                                 const int32_t* dst_row, int64_t R, float* const* dsts, int32_t n_dsts, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (R <= 0) return 0;
  if (!x || !src_local || !dst_sel || !dst_row || !dsts) { set_error("admm_pack_rows_to: null pointer"); return 1; }
  if (n_dsts < 1 || n_dsts > ADMM_MAX_DESTS) { set_error("admm_pack_rows_to: 1..%d destination buffers", ADMM_MAX_DESTS); return 1; }
  PackArgs a;
  for (int g = 0; g < 6; g++) {
    a.x[g] = x->g[g];
    if (!a.x[g]) { set_error("admm_pack_rows_to: null tensor"); return 1; }
  }
  for (int k = 0; k < ADMM_MAX_DESTS; k++) a.dst[k] = k < n_dsts ? dsts[k] : nullptr;
  admm_pack_rows_kernel<<<consensus_grid(R), ADMM_THREADS, 0, s>>>(a, src_local, dst_sel, dst_row, R, nullptr);
  return check_launch("admm_pack_rows_to", false, s);
}

extern "C" int admm_consensus_update(const AdmmConstTensors* x, const AdmmTensors* u, const AdmmTensors* z,
                                     const AdmmTensors* zmu, const int32_t* shared_row, const uint32_t* addends,
                                     const uint8_t* count, int32_t max_addends, int64_t n,
                                     const float* const* sources, int32_t n_sources, float relax,
                                     void* scratch, double* residuals, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (n <= 0) return 0;
  if (!x || !u || !z || !shared_row || !residuals || !scratch) { set_error("admm_consensus_update: null pointer"); return 1; }
  if (n_sources < 0 || n_sources > ADMM_MAX_SOURCES - 1) { set_error("admm_consensus_update: at most %d exchange buffers", ADMM_MAX_SOURCES - 1); return 1; }
  if (max_addends > 0 && (!addends || !count)) { set_error("admm_consensus_update: null addend table"); return 1; }
  ConsensusArgs a;
  for (int g = 0; g < 6; g++) {
    a.x[g] = x->g[g]; a.u[g] = u->g[g]; a.z[g] = z->g[g]; a.zmu[g] = zmu ? zmu->g[g] : nullptr;
    if (!a.x[g] || !a.u[g] || !a.z[g]) { set_error("admm_consensus_update: null tensor"); return 1; }
  }
  for (int k = 0; k < ADMM_MAX_SOURCES; k++) a.src[k] = (sources && k < n_sources) ? sources[k] : nullptr;
  double* partial = reinterpret_cast<double*>(scratch);
  unsigned int* ticket = reinterpret_cast<unsigned int*>(reinterpret_cast<char*>(scratch) + (size_t)(148 * 8) * 12 * sizeof(double));
  if (max_addends <= 4)
    admm_consensus_kernel<4><<<consensus_grid(n), ADMM_THREADS, 0, s>>>(a, shared_row, addends, count, max_addends, n,
                                                                       relax, partial, ticket, residuals);
  else
    admm_consensus_kernel<0><<<consensus_grid(n), ADMM_THREADS, 0, s>>>(a, shared_row, addends, count, max_addends, n,
                                                                       relax, partial, ticket, residuals);
  return check_launch("admm_consensus_update", false, s);
}

extern "C" int admm_penalty(const AdmmConstTensors* x, const AdmmConstTensors* u, const AdmmConstTensors* z,
                            const AdmmTensors* grad, int64_t n, const float* coef, double* sums,
                            void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (n <= 0) return 0;
  if (!x || !z || !coef || !sums) { set_error("admm_penalty: null pointer"); return 1; }
  PenaltyArgs a;
  unsigned long long end = 0;
  for (int g = 0; g < 6; g++) {
    a.x[g] = x->g[g]; a.u[g] = u ? u->g[g] : nullptr; a.z[g] = z->g[g];
    a.grad[g] = grad ? grad->g[g] : nullptr;
    a.len[g] = a.x[g] ? (unsigned long long)n * kSize[g] : 0ull;   // NULL x tensor: group handled elsewhere
    end += (a.len[g] + 3) / 4;
    a.chunk_end[g] = end;
    a.coef[g] = coef[g];
    const uintptr_t al = (uintptr_t)a.x[g] | (uintptr_t)a.u[g] | (uintptr_t)a.z[g] | (uintptr_t)a.grad[g];
    if (al & 15) { set_error("admm_penalty: tensors must be 16-byte aligned"); return 1; }
  }
  admm_penalty_kernel<<<grid_for(end), ADMM_THREADS, 0, s>>>(a, sums);
  return check_launch("admm_penalty", false, s);
}

// ---- block membership (re-split after fusion) ------------------------------------------------------
// select_gaussians_in_each_block / fuse_block_gaussians (master_gaussian_trainer.py:37-172) test every
// Gaussian against every block's 2-D bounding box on the host (numpy, one argwhere per block, after
// trimesh.transform_points with the world->OBB matrix).  One pass here: bit b of mask[i] = Gaussian i
// lies inside box b (closed interval, as conerf/datasets/utils.py:186-205), count[i] = popcount.
// The transform and the comparisons run in FP64 like numpy's (float32 points promoted to double).
namespace dogs {
struct BoxSet { double t[9]; int has_t; int n; };
__global__ void __launch_bounds__(256)
block_membership_kernel(int64_t n, const float* __restrict__ xyz, const BoxSet bs,
                        const double* __restrict__ boxes, unsigned long long* __restrict__ mask,
                        int32_t* __restrict__ count) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double x = (double)xyz[3 * i], y = (double)xyz[3 * i + 1];
  if (bs.has_t) {   // trimesh.transform_points on 2-D points: (M . [x, y, 1]^T)[:2], no perspective divide
    const double ox = bs.t[0] * x + bs.t[1] * y + bs.t[2];
    const double oy = bs.t[3] * x + bs.t[4] * y + bs.t[5];
    x = ox; y = oy;
  }
  unsigned long long m = 0ull;
  for (int b = 0; b < bs.n; b++) {
    const double ax = boxes[4 * b], ay = boxes[4 * b + 1], bx = boxes[4 * b + 2], by = boxes[4 * b + 3];
    if (ax <= x && x <= bx && ay <= y && y <= by) m |= 1ull << b;
  }
  mask[i] = m;
  if (count) count[i] = __popcll(m);
}
}  // namespace dogs

extern "C" int admm_block_membership(int64_t n, const float* xyz, const double* world_to_obb,
                                     int32_t n_blocks, const double* boxes, uint64_t* mask,
                                     int32_t* count, void* stream) {
  using namespace dogs;
  if (n < 0 || n_blocks < 1 || n_blocks > 64) { set_error("admm_block_membership: 1..64 blocks supported"); return 1; }
  if (n == 0) return 0;
  if (!xyz || !boxes || !mask) { set_error("admm_block_membership: null pointer"); return 1; }
  BoxSet bs;
  bs.has_t = world_to_obb != nullptr;
  bs.n = n_blocks;
  for (int k = 0; k < 9; k++) bs.t[k] = world_to_obb ? world_to_obb[k] : 0.0;
  cudaStream_t s = (cudaStream_t)stream;
  block_membership_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(
      n, xyz, bs, boxes, reinterpret_cast<unsigned long long*>(mask), count);
  return check_launch("admm_block_membership", false, s);
}

