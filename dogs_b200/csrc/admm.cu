// ADMM Gaussian-consensus kernels.  They replace the torch tensor arithmetic of
//   master_gaussian_trainer.py:538-555 (consensus average), :396-456 (primal/dual residuals),
//   slave_gaussian_trainer.py:99-121 (dual update) and :161-202 (penalty terms)
// for one scene block per GPU.  Only Gaussians owned by more than one block travel: they are
// packed into a dense [S_total, 59] buffer, summed across ranks with ONE NCCL all-reduce, and
// unpacked by kernels that also apply the dual update and accumulate the residual sums, so the
// 59 floats of a Gaussian are read and written once per consensus round.
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

// sendbuf[slot[i]*59 + START + off] = x[local_idx[i]*SZ + off], flat over the S*SZ elements of one group
template <int SZ>
__global__ void __launch_bounds__(ADMM_THREADS)
admm_pack_kernel(const float* __restrict__ x, const int64_t* __restrict__ local_idx,
                 const int64_t* __restrict__ slot, unsigned long long total, int start,
                 float* __restrict__ sendbuf) {
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    const unsigned long long i = e / SZ;
    const int off = (int)(e - i * SZ);
    sendbuf[slot[i] * ADMM_ROW + start + off] = x[local_idx[i] * SZ + off];
  }
}

// Over ALL n Gaussians of the block (flat over n*SZ elements of one group):
//   shared (slot >= 0): z = recv * inv_count (master:551-555); u += relax*(x - z) (slave:99-121)
//   single-owner      : z = x, u untouched
//   residuals[g] += sum (z - x)^2 (primal, master:396-432); residuals[6+g] += sum inv_count*(z_prev - z)^2 (dual, :435-456)
template <int SZ>
__global__ void __launch_bounds__(ADMM_THREADS)
admm_unpack_kernel(const float* __restrict__ x, float* __restrict__ u, float* __restrict__ z,
                   float* __restrict__ zmu, const int64_t* __restrict__ slot_of_local, const float* __restrict__ inv_count,
                   unsigned long long total, int start, const float* __restrict__ recvbuf, float relax,
                   double* __restrict__ residuals, int g) {
  __shared__ double acc[2];
  if (threadIdx.x < 2) acc[threadIdx.x] = 0.0;
  __syncthreads();
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  double primal = 0.0, dual = 0.0;
  for (unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    const unsigned long long i = e / SZ;
    const int off = (int)(e - i * SZ);
    const int64_t sl = slot_of_local[i];
    const float xv = x[e];
    const float z_old = z[e];
    float z_new;
    if (sl >= 0) {
      const float w = inv_count[i];
      z_new = recvbuf[sl * ADMM_ROW + start + off] * w;
      const float u_new = u[e] + relax * (xv - z_new);
      u[e] = u_new;
      if (zmu) zmu[e] = z_new - u_new;
      const float dp = z_new - xv;
      primal += (double)dp * dp;
      const float dd = z_old - z_new;
      dual += (double)w * ((double)dd * dd);
    } else {
      z_new = xv;
      const float dd = z_old - z_new;
      dual += (double)dd * dd;
      if (zmu) zmu[e] = z_new - u[e];
    }
    z[e] = z_new;
  }
  block_sum_to(primal, &acc[0]);
  block_sum_to(dual, &acc[1]);
  __syncthreads();
  if (threadIdx.x == 0) {
    if (acc[0] != 0.0) atomicAdd(&residuals[g], acc[0]);
    if (acc[1] != 0.0) atomicAdd(&residuals[6 + g], acc[1]);
  }
}

static const int kSize[6] = {3, 3, 45, 3, 4, 1};
static const int kStart[6] = {0, 3, 6, 51, 54, 58};

static int grid_for(unsigned long long work_items) {
  unsigned long long g = (work_items + ADMM_THREADS - 1) / ADMM_THREADS;
  const unsigned long long cap = 148ull * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace dogs

using namespace dogs;

extern "C" int admm_pack(const AdmmConstTensors* x, const int64_t* local_idx, const int64_t* slot,
                         int64_t S, int64_t S_total, int32_t zero_fill, float* sendbuf, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (!sendbuf && S_total > 0) { set_error("admm_pack: null send buffer"); return 1; }
  if (zero_fill && S_total > 0) cudaMemsetAsync(sendbuf, 0, (size_t)S_total * ADMM_ROW * sizeof(float), s);
  if (S <= 0) return check_launch("admm_pack", false, s, 0);
  if (!x || !local_idx || !slot) { set_error("admm_pack: null pointer"); return 1; }
#define PACK(G, SZ)                                                                                   \
  admm_pack_kernel<SZ><<<grid_for((unsigned long long)S * SZ), ADMM_THREADS, 0, s>>>(                 \
      x->g[G], local_idx, slot, (unsigned long long)S * SZ, kStart[G], sendbuf)
  PACK(0, 3); PACK(1, 3); PACK(2, 45); PACK(3, 3); PACK(4, 4); PACK(5, 1);
#undef PACK
  return check_launch("admm_pack", false, s, 6);
}

extern "C" int admm_unpack_update(const AdmmConstTensors* x, const AdmmTensors* u, const AdmmTensors* z,
                                  const AdmmTensors* zmu, const int64_t* slot_of_local, const float* inv_count, int64_t n,
                                  const float* recvbuf, float relax, double* residuals, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (n <= 0) return 0;
  if (!x || !u || !z || !slot_of_local || !inv_count || !residuals) { set_error("admm_unpack_update: null pointer"); return 1; }
#define UNPACK(G, SZ)                                                                                 \
  admm_unpack_kernel<SZ><<<grid_for((unsigned long long)n * SZ), ADMM_THREADS, 0, s>>>(               \
      x->g[G], u->g[G], z->g[G], zmu ? zmu->g[G] : nullptr, slot_of_local, inv_count, (unsigned long long)n * SZ, kStart[G], recvbuf, \
      relax, residuals, G)
  UNPACK(0, 3); UNPACK(1, 3); UNPACK(2, 45); UNPACK(3, 3); UNPACK(4, 4); UNPACK(5, 1);
#undef UNPACK
  return check_launch("admm_unpack_update", false, s, 6);
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

