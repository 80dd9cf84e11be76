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
//
// Work layout (second version).  The first one gave a warp one Gaussian: its 59 floats sit in six tensors, so
// every load or store instruction of the warp touched up to six different cache lines and used 12-180 bytes of
// each — the kernel ran at 35 % of the HBM peak, bound by the L1 tag stage, not by DRAM.  Now a CTA takes a tile
// of 256 consecutive Gaussians, stages their table entries (shared row, owner count, addend entries) in shared
// memory once, and then walks the tile TENSOR BY TENSOR: the tile's slice of a tensor is one contiguous run
// (256 x 3 ... 256 x 45 floats), consecutive threads take consecutive floats, four floats per thread are in
// flight before anything is summed.  Addend rows come from the 59-float exchange rows; the floats of one group
// are contiguous there as well.
constexpr int CT_ROWS = 256;          // Gaussians per tile
#ifndef CT_MIN_CTAS
#define CT_MIN_CTAS 4
#endif
constexpr uint32_t ENT_NONE = 0xffffffffu;

template <int SZ>
__device__ __noinline__ void consensus_group_tail(const ConsensusArgs& a, int g, int row_off, int64_t i0, int rows,
                                                  const int32_t* s_sr, const uint8_t* s_cnt, const uint32_t* s_ent,
                                                  int max_addends, float relax, bool has_zmu, const double* inv_cnt,
                                                  double* s_acc);
// One element: average / dual update / residual terms (shared by the vector and the tail path).
__device__ __forceinline__ void consensus_element(const ConsensusArgs& a, int row_off, int c, int32_t sr, int cnt_i,
                                                  const uint32_t* ent, int max_addends, float relax, float xv, float zo,
                                                  float uv, const double* inv_cnt, float& z_new, float& u_new,
                                                  double& primal, double& dual) {
  z_new = xv;
  u_new = uv;
  if (sr >= 0) {
    float sum = 0.f;
    for (int j = 0; j < max_addends; j++) {
      const uint32_t en = ent[j];
      if (en == ENT_NONE) break;
      const uint32_t which = en >> 28;
      const float v = which == 15u ? xv : __ldg(a.src[which] + (int64_t)(en & 0x0fffffffu) * ADMM_ROW + row_off + c);
      sum = j == 0 ? v : __fadd_rn(sum, v);
    }
    z_new = __fdiv_rn(sum, (float)cnt_i);
    u_new = __fadd_rn(uv, __fmul_rn(relax, __fsub_rn(xv, z_new)));
  }
  const double dp = (double)z_new - (double)xv, dd = (double)zo - (double)z_new;
  primal += dp * dp;
  dual += dd * dd * inv_cnt[cnt_i];
}

template <int SZ>
__device__ __noinline__ void consensus_group_tail(const ConsensusArgs& a, int g, int row_off, int64_t i0, int rows,
                                                  const int32_t* s_sr, const uint8_t* s_cnt, const uint32_t* s_ent,
                                                  int max_addends, float relax, bool has_zmu, const double* inv_cnt,
                                                  double* s_acc) {
  double primal = 0.0, dual = 0.0;
  const float* xg = a.x[g] + i0 * SZ;
  float* ug = a.u[g] + i0 * SZ;
  float* zg = a.z[g] + i0 * SZ;
  float* mg = has_zmu ? a.zmu[g] + i0 * SZ : nullptr;
  for (int e = threadIdx.x; e < rows * SZ; e += ADMM_THREADS) {
    const int r = e / SZ, c = e - r * SZ;
    const int32_t sr = s_sr[r];
    float z_new, u_new;
    consensus_element(a, row_off, c, sr, sr >= 0 ? (int)s_cnt[r] : 1, s_ent + r * max_addends, max_addends, relax, xg[e],
                      zg[e], ug[e], inv_cnt, z_new, u_new, primal, dual);
    zg[e] = z_new;
    if (sr >= 0) ug[e] = u_new;
    if (has_zmu) mg[e] = z_new - u_new;
  }
  s_acc[g * ADMM_THREADS + threadIdx.x] += primal;
  s_acc[(6 + g) * ADMM_THREADS + threadIdx.x] += dual;
}

// A full tile of one tensor as 16-byte vectors: 256 x SZ floats, contiguous and 16-byte aligned.
// MAXA > 0: at most MAXA addends per Gaussian; the (up to) 4 x MAXA addend floats of a vector are all requested
// before anything is summed — taken one after the other (load, add, next entry, load ...) the gathers formed a
// chain of up to 12 memory round trips per vector.  MAXA = 0: any number of addends, sequential.
template <int SZ, int MAXA>
__device__ __forceinline__ void consensus_group(const ConsensusArgs& a, int g, int row_off, int64_t i0,
                                                const int32_t* s_sr, const uint8_t* s_cnt, const uint32_t* s_ent,
                                                int max_addends, float relax, bool has_zmu, const double* inv_cnt,
                                                double* s_acc) {
  const float4* __restrict__ x4 = reinterpret_cast<const float4*>(a.x[g] + i0 * SZ);
  float4* __restrict__ u4 = reinterpret_cast<float4*>(a.u[g] + i0 * SZ);
  float4* __restrict__ z4 = reinterpret_cast<float4*>(a.z[g] + i0 * SZ);
  float4* __restrict__ m4 = has_zmu ? reinterpret_cast<float4*>(a.zmu[g] + i0 * SZ) : nullptr;
  constexpr int total4 = CT_ROWS * SZ / 4;
  double primal = 0.0, dual = 0.0;
  for (int q = threadIdx.x; q < total4; q += ADMM_THREADS) {
    const float4 xv = x4[q], zo = z4[q], uv = u4[q];
    const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, zs[4] = {zo.x, zo.y, zo.z, zo.w}, us[4] = {uv.x, uv.y, uv.z, uv.w};
    float zn[4], un[4];
    bool any_shared = false;
    if (MAXA > 0) {
      float add[4][MAXA > 0 ? MAXA : 1];
      int rr[4];
#pragma unroll
      for (int c4 = 0; c4 < 4; c4++) {
        const int e = 4 * q + c4, r = e / SZ, c = e - r * SZ;
        rr[c4] = r;
        const bool sh = s_sr[r] >= 0;
        any_shared |= sh;
#pragma unroll
        for (int j = 0; j < MAXA; j++) {
          const uint32_t en = (sh && j < max_addends) ? s_ent[r * max_addends + j] : ENT_NONE;
          const uint32_t which = en >> 28;
          add[c4][j] = (en == ENT_NONE || which == 15u)
                           ? 0.f
                           : __ldg(a.src[which] + (int64_t)(en & 0x0fffffffu) * ADMM_ROW + row_off + c);
        }
      }
#pragma unroll
      for (int c4 = 0; c4 < 4; c4++) {
        const int r = rr[c4];
        const bool sh = s_sr[r] >= 0;
        const int cnt_i = sh ? (int)s_cnt[r] : 1;
        float z_new = xs[c4], u_new = us[c4];
        if (sh) {
          float sum = 0.f;
#pragma unroll
          for (int j = 0; j < MAXA; j++) {
            const uint32_t en = j < max_addends ? s_ent[r * max_addends + j] : ENT_NONE;
            if (en != ENT_NONE) {                 // (entries are packed: nothing follows the first empty one)
              const float v = (en >> 28) == 15u ? xs[c4] : add[c4][j];
              sum = j == 0 ? v : __fadd_rn(sum, v);
            }
          }
          z_new = __fdiv_rn(sum, (float)cnt_i);
          u_new = __fadd_rn(us[c4], __fmul_rn(relax, __fsub_rn(xs[c4], z_new)));
        }
        zn[c4] = z_new; un[c4] = u_new;
        const double dp = (double)z_new - (double)xs[c4], dd = (double)zs[c4] - (double)z_new;
        primal += dp * dp;
        dual += dd * dd * inv_cnt[cnt_i];
      }
    } else {
#pragma unroll
      for (int c4 = 0; c4 < 4; c4++) {
        const int e = 4 * q + c4, r = e / SZ, c = e - r * SZ;
        const int32_t sr = s_sr[r];
        any_shared |= sr >= 0;
        consensus_element(a, row_off, c, sr, sr >= 0 ? (int)s_cnt[r] : 1, s_ent + r * max_addends, max_addends, relax,
                          xs[c4], zs[c4], us[c4], inv_cnt, zn[c4], un[c4], primal, dual);
      }
    }
    z4[q] = make_float4(zn[0], zn[1], zn[2], zn[3]);
    if (any_shared) u4[q] = make_float4(un[0], un[1], un[2], un[3]);     // single owners: u is unchanged
    if (has_zmu) m4[q] = make_float4(zn[0] - un[0], zn[1] - un[1], zn[2] - un[2], zn[3] - un[3]);
  }
  // per-thread, per-group running sums live in shared memory between tiles (12 doubles per thread would cost
  // 24 registers of a kernel that needs its occupancy)
  s_acc[g * ADMM_THREADS + threadIdx.x] += primal;
  s_acc[(6 + g) * ADMM_THREADS + threadIdx.x] += dual;
}

template <int MAXA>
__global__ void __launch_bounds__(ADMM_THREADS, CT_MIN_CTAS)
admm_consensus_kernel(const __grid_constant__ ConsensusArgs a, const int32_t* __restrict__ shared_row,
                      const uint32_t* __restrict__ addends, const uint8_t* __restrict__ count, int max_addends,
                      int64_t n, float relax, double* __restrict__ partial, unsigned int* __restrict__ ticket,
                      double* __restrict__ residuals) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* s_acc = reinterpret_cast<double*>(smem_raw);                                  // [12][ADMM_THREADS]
  uint32_t* s_ent = reinterpret_cast<uint32_t*>(smem_raw + 12 * ADMM_THREADS * sizeof(double));   // [CT_ROWS][max_addends]
  __shared__ int32_t s_sr[CT_ROWS];
  __shared__ uint8_t s_cnt[CT_ROWS];
  __shared__ double inv_cnt[256];              // 1 / count without a double division per element
  __shared__ bool last;
#pragma unroll
  for (int q = 0; q < 12; q++) s_acc[q * ADMM_THREADS + threadIdx.x] = 0.0;
  inv_cnt[threadIdx.x] = 1.0 / (double)max((int)threadIdx.x, 1);
  const bool has_zmu = a.zmu[0] != nullptr;
  const int64_t ntiles = (n + CT_ROWS - 1) / CT_ROWS;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t i0 = tile * CT_ROWS;
    const int rows = (int)min((int64_t)CT_ROWS, n - i0);
    __syncthreads();                           // the previous tile is done with the tables (and inv_cnt is set)
    if ((int)threadIdx.x < rows) {
      const int32_t sr = shared_row[i0 + threadIdx.x];
      s_sr[threadIdx.x] = sr;
      s_cnt[threadIdx.x] = sr >= 0 ? count[sr] : (uint8_t)1;
      const uint32_t* src = addends + (int64_t)(sr >= 0 ? sr : 0) * max_addends;
      for (int j = 0; j < max_addends; j++) s_ent[threadIdx.x * max_addends + j] = sr >= 0 ? src[j] : ENT_NONE;
    }
    __syncthreads();
    if (rows == CT_ROWS) {
      consensus_group<3, MAXA>(a, 0, 0, i0, s_sr, s_cnt, s_ent, max_addends, relax, has_zmu, inv_cnt, s_acc);
      consensus_group<3, MAXA>(a, 1, 3, i0, s_sr, s_cnt, s_ent, max_addends, relax, has_zmu, inv_cnt, s_acc);
      consensus_group<45, MAXA>(a, 2, 6, i0, s_sr, s_cnt, s_ent, max_addends, relax, has_zmu, inv_cnt, s_acc);
      consensus_group<3, MAXA>(a, 3, 51, i0, s_sr, s_cnt, s_ent, max_addends, relax, has_zmu, inv_cnt, s_acc);
      consensus_group<4, MAXA>(a, 4, 54, i0, s_sr, s_cnt, s_ent, max_addends, relax, has_zmu, inv_cnt, s_acc);
      consensus_group<1, MAXA>(a, 5, 58, i0, s_sr, s_cnt, s_ent, max_addends, relax, has_zmu, inv_cnt, s_acc);
    } else {      // the last, partial tile: element by element
      consensus_group_tail<3>(a, 0, 0, i0, rows, s_sr, s_cnt, s_ent, max_addends, relax, has_zmu, inv_cnt, s_acc);
      consensus_group_tail<3>(a, 1, 3, i0, rows, s_sr, s_cnt, s_ent, max_addends, relax, has_zmu, inv_cnt, s_acc);
      consensus_group_tail<45>(a, 2, 6, i0, rows, s_sr, s_cnt, s_ent, max_addends, relax, has_zmu, inv_cnt, s_acc);
      consensus_group_tail<3>(a, 3, 51, i0, rows, s_sr, s_cnt, s_ent, max_addends, relax, has_zmu, inv_cnt, s_acc);
      consensus_group_tail<4>(a, 4, 54, i0, rows, s_sr, s_cnt, s_ent, max_addends, relax, has_zmu, inv_cnt, s_acc);
      consensus_group_tail<1>(a, 5, 58, i0, rows, s_sr, s_cnt, s_ent, max_addends, relax, has_zmu, inv_cnt, s_acc);
    }
  }
  __syncthreads();
  // fold the per-thread sums: warp w reduces sums w, w + 8 (12 sums, 8 warps)
  __shared__ double acc[12];
  for (int q = threadIdx.x >> 5; q < 12; q += ADMM_THREADS / 32) {
    double v = 0.0;
    for (int t = threadIdx.x & 31; t < ADMM_THREADS; t += 32) v += s_acc[q * ADMM_THREADS + t];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) acc[q] = v;
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
  const int64_t ntiles = (n + CT_ROWS - 1) / CT_ROWS;
  const int grid = (int)(ntiles < 148 * 8 ? ntiles : 148 * 8);
  const size_t smem = 12 * ADMM_THREADS * sizeof(double) + (size_t)CT_ROWS * (max_addends > 0 ? max_addends : 1) * sizeof(uint32_t);
  if (max_addends > 16) { set_error("admm_consensus_update: more than 16 addends per Gaussian"); return 1; }
  if (max_addends <= 4)
    admm_consensus_kernel<4><<<grid, ADMM_THREADS, smem, s>>>(a, shared_row, addends, count, max_addends, n, relax, partial,
                                                             ticket, residuals);
  else
    admm_consensus_kernel<0><<<grid, ADMM_THREADS, smem, s>>>(a, shared_row, addends, count, max_addends, n, relax, partial,
                                                             ticket, residuals);
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

