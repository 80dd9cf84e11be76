// ADMM Gaussian-consensus kernels.  They replace the torch tensor arithmetic of
//   master_gaussian_trainer.py:538-555 (consensus average), :396-456 (primal/dual residuals),
//   slave_gaussian_trainer.py:99-121 (dual update) and :161-202 (penalty terms)
// for one scene block per GPU.  Only Gaussians owned by more than one block travel: they are
// packed into a dense [S_total, 59] buffer, summed across ranks with ONE NCCL all-reduce, and
// unpacked by a kernel that also applies the dual update and accumulates the residual sums, so
// the 59 floats of a Gaussian are read and written once per consensus round.
//
// A Gaussian = 59 floats in 6 tensors: xyz 3, f_dc 3, f_rest 45, scaling 3, quaternion 4, opacity 1
// (raw, pre-activation values are the consensus variables).
#include "common.cuh"

namespace dogs {

constexpr int ADMM_ROW = 59;
__device__ __forceinline__ int admm_group_of(int e) {
  return e < 3 ? 0 : e < 6 ? 1 : e < 51 ? 2 : e < 54 ? 3 : e < 58 ? 4 : e < 59 ? 5 : -1;
}
__constant__ int ADMM_START[6] = {0, 3, 6, 51, 54, 58};
__constant__ int ADMM_SIZE[6] = {3, 3, 45, 3, 4, 1};

struct Ptrs6 { float* g[6]; };
struct CPtrs6 { const float* g[6]; };

// blockDim = (64, 4): x = element of the 59-float row, y = Gaussian within the CTA.
__global__ void __launch_bounds__(256)
admm_pack_kernel(CPtrs6 x, const int64_t* __restrict__ local_idx, const int64_t* __restrict__ slot,
                 int64_t S, float* __restrict__ sendbuf) {
  const int e = threadIdx.x;
  const int g = admm_group_of(e);
  if (g < 0) return;
  const int off = e - ADMM_START[g], sz = ADMM_SIZE[g];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.y + threadIdx.y; i < S; i += (int64_t)gridDim.x * blockDim.y)
    sendbuf[slot[i] * ADMM_ROW + e] = x.g[g][local_idx[i] * sz + off];
}

__global__ void __launch_bounds__(256)
admm_unpack_update_kernel(CPtrs6 x, Ptrs6 u, Ptrs6 z, const int64_t* __restrict__ slot_of_local,
                          const float* __restrict__ inv_count, int64_t n,
                          const float* __restrict__ recvbuf, float relax, double* __restrict__ residuals) {
  __shared__ double acc[12];
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  if (tid < 12) acc[tid] = 0.0;
  __syncthreads();
  const int e = threadIdx.x;
  const int g = admm_group_of(e);
  double primal = 0.0, dual = 0.0;
  if (g >= 0) {
    const int off = e - ADMM_START[g], sz = ADMM_SIZE[g];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.y + threadIdx.y; i < n; i += (int64_t)gridDim.x * blockDim.y) {
      const int64_t k = i * sz + off;
      const int64_t sl = slot_of_local[i];
      const float xv = x.g[g][k];
      const float z_old = z.g[g][k];
      float z_new;
      if (sl >= 0) {
        const float w = inv_count[i];
        z_new = recvbuf[sl * ADMM_ROW + e] * w;            // master:551-555 average_gaussians
        u.g[g][k] += relax * (xv - z_new);                 // slave:99-121
        const float dp = z_new - xv;
        primal += (double)dp * dp;                         // master:396-432
        const float dd = z_old - z_new;
        dual += (double)w * ((double)dd * dd);             // master:435-456 (each global Gaussian once)
      } else {
        z_new = xv;                                        // count == 1: z == x, u unchanged
        const float dd = z_old - z_new;
        dual += (double)dd * dd;
      }
      z.g[g][k] = z_new;
    }
  }
  if (g >= 0) {
    atomicAdd(&acc[g], primal);
    atomicAdd(&acc[6 + g], dual);
  }
  __syncthreads();
  if (tid < 12 && acc[tid] != 0.0) atomicAdd(&residuals[tid], acc[tid]);
}

__global__ void __launch_bounds__(256)
admm_penalty_kernel(CPtrs6 x, CPtrs6 u, CPtrs6 z, Ptrs6 grad, int64_t n, float c0, float c1, float c2,
                    float c3, float c4, float c5, double* __restrict__ sums) {
  __shared__ double acc[6];
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  if (tid < 6) acc[tid] = 0.0;
  __syncthreads();
  const int e = threadIdx.x;
  const int g = admm_group_of(e);
  double s = 0.0;
  if (g >= 0) {
    const float coef = g == 0 ? c0 : g == 1 ? c1 : g == 2 ? c2 : g == 3 ? c3 : g == 4 ? c4 : c5;
    const int off = e - ADMM_START[g], sz = ADMM_SIZE[g];
    float* gr = grad.g[g];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.y + threadIdx.y; i < n; i += (int64_t)gridDim.x * blockDim.y) {
      const int64_t k = i * sz + off;
      const float r = (x.g[g][k] + u.g[g][k]) - z.g[g][k];   // mse_loss(x + u, z), slave:169-187
      s += (double)r * r;
      if (gr) gr[k] += coef * r;                              // d/dx 0.5*rho*mean(r^2) = rho/numel * r
    }
    atomicAdd(&acc[g], s);
  }
  __syncthreads();
  if (tid < 6 && acc[tid] != 0.0) atomicAdd(&sums[tid], acc[tid]);
}

static int grid_for(int64_t rows) {
  int64_t g = (rows + 3) / 4;
  const int64_t cap = 148 * 8;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace dogs

using namespace dogs;

extern "C" int admm_pack(const AdmmConstTensors* x, const int64_t* local_idx, const int64_t* slot,
                         int64_t S, int64_t S_total, int32_t zero_fill, float* sendbuf, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (!sendbuf && S_total > 0) { set_error("admm_pack: null send buffer"); return 1; }
  if (zero_fill && S_total > 0) cudaMemsetAsync(sendbuf, 0, (size_t)S_total * ADMM_ROW * sizeof(float), s);
  if (S <= 0) return check_launch("admm_pack", false, s);
  if (!x || !local_idx || !slot) { set_error("admm_pack: null pointer"); return 1; }
  CPtrs6 xp;
  for (int g = 0; g < 6; g++) xp.g[g] = x->g[g];
  admm_pack_kernel<<<grid_for(S), dim3(64, 4), 0, s>>>(xp, local_idx, slot, S, sendbuf);
  return check_launch("admm_pack", false, s);
}

extern "C" int admm_unpack_update(const AdmmConstTensors* x, const AdmmTensors* u, const AdmmTensors* z,
                                  const int64_t* slot_of_local, const float* inv_count, int64_t n,
                                  const float* recvbuf, float relax, double* residuals, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (n <= 0) return 0;
  if (!x || !u || !z || !slot_of_local || !inv_count || !residuals) { set_error("admm_unpack_update: null pointer"); return 1; }
  CPtrs6 xp; Ptrs6 up, zp;
  for (int g = 0; g < 6; g++) { xp.g[g] = x->g[g]; up.g[g] = u->g[g]; zp.g[g] = z->g[g]; }
  admm_unpack_update_kernel<<<grid_for(n), dim3(64, 4), 0, s>>>(xp, up, zp, slot_of_local, inv_count, n,
                                                                recvbuf, relax, residuals);
  return check_launch("admm_unpack_update", false, s);
}

extern "C" int admm_penalty(const AdmmConstTensors* x, const AdmmConstTensors* u, const AdmmConstTensors* z,
                            const AdmmTensors* grad, int64_t n, const float* coef, double* sums,
                            void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (n <= 0) return 0;
  if (!x || !u || !z || !coef || !sums) { set_error("admm_penalty: null pointer"); return 1; }
  CPtrs6 xp, up, zp; Ptrs6 gp;
  for (int g = 0; g < 6; g++) {
    xp.g[g] = x->g[g]; up.g[g] = u->g[g]; zp.g[g] = z->g[g];
    gp.g[g] = grad ? grad->g[g] : nullptr;
  }
  admm_penalty_kernel<<<grid_for(n), dim3(64, 4), 0, s>>>(xp, up, zp, gp, n, coef[0], coef[1], coef[2],
                                                          coef[3], coef[4], coef[5], sums);
  return check_launch("admm_penalty", false, s);
}
