// Densification / pruning / opacity reset of the Gaussian SoA, replacing
// GaussianSplatModel.densify_and_prune / densify_and_clone / densify_and_split / prune_points /
// densification_postfix / reset_opacity / add_densification_stats
// (conerf/model/gaussian_fields/gaussian_splat_model.py:362-541) and the optimizer surgery they call
// (cat_tensors_to_optimizer / prune_optimizer / replace_tensor_to_optimizer, :34-108).
//
// The reference runs densify_and_prune as ~60 torch kernels: boolean masks, torch.cat of 6 parameter and 12
// Adam-state tensors (twice: clone, then split), a prune of the split parents, a second prune by opacity /
// size — four full copies of the model.  The outcome is a pure function of the inputs, and the order of the
// surviving rows is fixed:
//      [originals that were not split] [clones] [split replicas 0] [split replicas 1],
// each segment in ascending source order, minus the rows the final prune removes.  Here one kernel classifies
// every Gaussian (clone / split), one decides for each of its four possible output rows whether it survives,
// ONE exclusive scan over the 4 P flags (segment-major) yields every destination, and each tensor is then
// written once by a gather.
#include "common.cuh"

namespace dogs {

// max over the three activated scales, as torch.max(get_scaling, dim=1) sees it
__device__ __forceinline__ float max_scale(const float* __restrict__ raw_scaling, int64_t i) {
  return fmaxf(fmaxf(expf(raw_scaling[3 * i]), expf(raw_scaling[3 * i + 1])), expf(raw_scaling[3 * i + 2]));
}

// densify_and_prune :502-506 (grads = accum / denom, NaN -> 0), densify_and_clone :436-443,
// densify_and_split :462-470.  cls bit 0 = clone, bit 1 = split; split_flag[i] = 1/0 for the rank scan.
__global__ void __launch_bounds__(256)
densify_classify_kernel(int64_t P, const float* __restrict__ grad_accum, const float* __restrict__ denom,
                        const float* __restrict__ raw_scaling, float grad_threshold, float scale_threshold,
                        uint8_t* __restrict__ cls, uint32_t* __restrict__ split_flag) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  float g = __fdiv_rn(grad_accum[i], denom[i]);
  if (isnan(g)) g = 0.f;
  // torch.norm(grads, dim=-1) of a [P,1] tensor is |g|; the split test uses g itself
  const float ms = max_scale(raw_scaling, i);
  const bool clone = fabsf(g) >= grad_threshold && ms <= scale_threshold;
  const bool split = g >= grad_threshold && ms > scale_threshold;
  cls[i] = (uint8_t)((clone ? 1 : 0) | (split ? 2 : 0));
  split_flag[i] = split ? 1u : 0u;
}

// quaternion_to_rotation_mat (conerf/model/gaussian_fields/utils.py:27-70) applied to one sample:
// new_xyz = R(q / |q|) * sample + xyz   (densify_and_split :472-481)
__device__ __forceinline__ float3 split_position(const float* __restrict__ xyz, const float* __restrict__ quat,
                                                 const float* __restrict__ sample, int64_t i) {
  const float qr = quat[4 * i], qx = quat[4 * i + 1], qy = quat[4 * i + 2], qz = quat[4 * i + 3];
  const float n = __fsqrt_rn(__fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(qr, qr), __fmul_rn(qx, qx)), __fmul_rn(qy, qy)), __fmul_rn(qz, qz)));
  const float r = __fdiv_rn(qr, n), x = __fdiv_rn(qx, n), y = __fdiv_rn(qy, n), z = __fdiv_rn(qz, n);
  const float sx = sample[0], sy = sample[1], sz = sample[2];
  const float R00 = 1.f - 2.f * (y * y + z * z), R01 = 2.f * (x * y - r * z), R02 = 2.f * (x * z + r * y);
  const float R10 = 2.f * (x * y + r * z), R11 = 1.f - 2.f * (x * x + z * z), R12 = 2.f * (y * z - r * x);
  const float R20 = 2.f * (x * z - r * y), R21 = 2.f * (y * z + r * x), R22 = 1.f - 2.f * (x * x + y * y);
  return {R00 * sx + R01 * sy + R02 * sz + xyz[3 * i], R10 * sx + R11 * sy + R12 * sz + xyz[3 * i + 1],
          R20 * sx + R21 * sy + R22 * sz + xyz[3 * i + 2]};
}

// Survival of the four possible output rows of source Gaussian i under the final prune of densify_and_prune
// (:511-527): opacity < min_opacity, z below the bounding box, world-space size > 0.1 * extent.  (The screen-size
// test `max_radii2D > max_screen_size` never fires in the reference: densification_postfix :402-404 has just
// zeroed max_radii2D for every row.)  keep[seg * P + i], seg 0 original, 1 clone, 2 / 3 split replicas.
__global__ void __launch_bounds__(256)
densify_plan_kernel(int64_t P, const uint8_t* __restrict__ cls, const uint32_t* __restrict__ split_rank,
                    int64_t n_split, const float* __restrict__ xyz, const float* __restrict__ raw_scaling,
                    const float* __restrict__ raw_quat, const float* __restrict__ raw_opacity,
                    const float* __restrict__ samples, float min_opacity, float big_ws /* < 0: no size test */,
                    int has_zmin, float zmin, int num_replica, uint32_t* __restrict__ keep) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  const uint8_t c = cls[i];
  const float op = act_sigmoid(raw_opacity[i]);
  const float ms = max_scale(raw_scaling, i);
  bool pruned = op < min_opacity;
  if (has_zmin) pruned |= xyz[3 * i + 2] < zmin;
  if (big_ws >= 0.f) pruned |= ms > big_ws;
  keep[i] = (!(c & 2) && !pruned) ? 1u : 0u;               // split parents are removed (:492-497)
  keep[P + i] = ((c & 1) && !pruned) ? 1u : 0u;            // a clone is an exact copy: same verdict
  for (int r = 0; r < 2; r++) {
    bool k = false;
    if ((c & 2) && r < num_replica) {
      bool pr = op < min_opacity;
      if (has_zmin) {
        const float3 p = split_position(xyz, raw_quat, samples + 3 * ((int64_t)r * n_split + split_rank[i]), i);
        pr |= p.z < zmin;
      }
      // new scaling = log(exp(s) / (0.8 * num_replica)): the activated value is exp(s) / 1.6 again (up to the
      // rounding of log and exp); test what the reference tests: exp(new_raw_scaling)
      if (big_ws >= 0.f) {
        const float div = 0.8f * (float)num_replica;
        float m = 0.f;
        for (int a = 0; a < 3; a++) m = fmaxf(m, expf(logf(__fdiv_rn(expf(raw_scaling[3 * i + a]), div))));
        pr |= m > big_ws;
      }
      k = !pr;
    }
    keep[(2 + r) * P + i] = k ? 1u : 0u;
  }
}

// out row j <- (source Gaussian, segment): src_idx[j] = i, kind[j] = seg
__global__ void __launch_bounds__(256)
densify_scatter_kernel(int64_t P, const uint32_t* __restrict__ keep, const uint32_t* __restrict__ dest,
                       int32_t* __restrict__ src_idx, uint8_t* __restrict__ kind) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= 4 * P) return;
  if (keep[e]) {
    const int64_t seg = e / P;
    src_idx[dest[e]] = (int32_t)(e - seg * P);
    kind[dest[e]] = (uint8_t)seg;
  }
}

// dst[j] = src[src_idx[j]], or 0 for new rows when zero_new (Adam moments of appended Gaussians,
// cat_tensors_to_optimizer :60-66).  Rows are row_words 32-bit words.
__global__ void __launch_bounds__(256)
gather_rows_kernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, uint32_t row,
                   const int32_t* __restrict__ src_idx, const uint8_t* __restrict__ kind, int zero_new,
                   unsigned long long total) {
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    const unsigned long long j = e / row;
    const uint32_t off = (uint32_t)(e - j * row);
    uint32_t v = 0u;
    if (!(zero_new && kind && kind[j] != 0)) v = src[(unsigned long long)src_idx[j] * row + off];
    dst[e] = v;
  }
}

// split replicas: position and scaling differ from the parent's (:472-485)
__global__ void __launch_bounds__(256)
densify_split_fixup_kernel(int64_t n_out, const int32_t* __restrict__ src_idx, const uint8_t* __restrict__ kind,
                           const uint32_t* __restrict__ split_rank, int64_t n_split, const float* __restrict__ xyz,
                           const float* __restrict__ raw_scaling, const float* __restrict__ raw_quat,
                           const float* __restrict__ samples, int num_replica, float* __restrict__ xyz_out,
                           float* __restrict__ scaling_out) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_out) return;
  const int k = kind[j];
  if (k < 2) return;
  const int64_t i = src_idx[j];
  const float3 p = split_position(xyz, raw_quat, samples + 3 * ((int64_t)(k - 2) * n_split + split_rank[i]), i);
  xyz_out[3 * j] = p.x; xyz_out[3 * j + 1] = p.y; xyz_out[3 * j + 2] = p.z;
  const float div = 0.8f * (float)num_replica;
  for (int a = 0; a < 3; a++) scaling_out[3 * j + a] = logf(__fdiv_rn(expf(raw_scaling[3 * i + a]), div));
}

// reset_opacity (:362-367): opacity = inverse_sigmoid(min(sigmoid(opacity), 0.01)), Adam moments of the tensor zeroed
// (replace_tensor_to_optimizer :34-48).
__global__ void __launch_bounds__(256)
reset_opacity_kernel(int64_t P, float* __restrict__ raw_opacity, float* __restrict__ exp_avg,
                     float* __restrict__ exp_avg_sq, float cap) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  const float x = fminf(act_sigmoid(raw_opacity[i]), cap);
  raw_opacity[i] = logf(__fdiv_rn(x, __fsub_rn(1.0f, x)));
  if (exp_avg) exp_avg[i] = 0.f;
  if (exp_avg_sq) exp_avg_sq[i] = 0.f;
}

// add_densification_stats (:529-541) + the max_radii2D update of the trainer (gaussian_trainer.py:433-437):
// for visible Gaussians: accum += |grad_means2D.xy|, denom += 1, max_radii2D = max(max_radii2D, radii).
__global__ void __launch_bounds__(256)
densify_stats_kernel(int64_t P, const float* __restrict__ grad_means2D, const int32_t* __restrict__ radii,
                     float* __restrict__ grad_accum, float* __restrict__ denom, float* __restrict__ max_radii2D) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  const int r = radii[i];
  if (r <= 0) return;
  const float gx = grad_means2D[3 * i], gy = grad_means2D[3 * i + 1];
  grad_accum[i] = __fadd_rn(grad_accum[i], __fsqrt_rn(__fadd_rn(__fmul_rn(gx, gx), __fmul_rn(gy, gy))));
  denom[i] = __fadd_rn(denom[i], 1.0f);
  if (max_radii2D) max_radii2D[i] = fmaxf(max_radii2D[i], (float)r);
}

}  // namespace dogs

using namespace dogs;

static unsigned grid1(int64_t n) { return (unsigned)((n + 255) / 256); }

extern "C" int dgr_densify_classify(int64_t P, const float* grad_accum, const float* denom, const float* raw_scaling,
                                    float grad_threshold, float scale_threshold, uint8_t* cls, uint32_t* split_rank,
                                    dogs_alloc_fn scratch_alloc, void* scratch_user, int64_t* n_split_out, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (!n_split_out) { set_error("densify_classify: null n_split_out"); return 1; }
  *n_split_out = 0;
  if (P <= 0) return 0;
  if (P > 0x3fffffffll) { set_error("densify_classify: too many Gaussians"); return 1; }
  if (!grad_accum || !denom || !raw_scaling || !cls || !split_rank || !scratch_alloc) { set_error("densify_classify: null pointer"); return 1; }
  const size_t n = (size_t)P;
  char* sc = scratch_alloc(scratch_user, align_up(n * 4) + scan_scratch_bytes(n) + 512);
  if (!sc) { set_error("densify_classify: scratch allocation failed"); return 1; }
  Carver cv(sc);
  uint32_t* flags = cv.take<uint32_t>(n);
  uint32_t* total = cv.take<uint32_t>(1);
  char* scan_scratch = cv.take<char>(scan_scratch_bytes(n));
  densify_classify_kernel<<<grid1(P), 256, 0, s>>>(P, grad_accum, denom, raw_scaling, grad_threshold, scale_threshold, cls, flags);
  if (exclusive_scan_u32(flags, nullptr, split_rank, total, n, scan_scratch, false, s)) return 1;
  uint32_t h = 0;
  cudaError_t e = cudaMemcpyAsync(&h, total, 4, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);      // the caller draws 2 * n_split normal samples next
  if (e != cudaSuccess) { set_error("[CUDA ERROR] densify_classify: %s", cudaGetErrorString(e)); return 1; }
  *n_split_out = (int64_t)h;
  return check_launch("densify_classify", false, s, 4);
}

extern "C" int dgr_densify_plan(int64_t P, const uint8_t* cls, const uint32_t* split_rank, int64_t n_split,
                                const float* xyz, const float* raw_scaling, const float* raw_quat,
                                const float* raw_opacity, const float* samples, float min_opacity, float big_ws,
                                int32_t has_zmin, float zmin, int32_t num_replica, int32_t* src_idx, uint8_t* kind,
                                dogs_alloc_fn scratch_alloc, void* scratch_user, int64_t* n_out, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (!n_out) { set_error("densify_plan: null n_out"); return 1; }
  *n_out = 0;
  if (P <= 0) return 0;
  if (num_replica < 1 || num_replica > 2) { set_error("densify_plan: num_replica must be 1 or 2"); return 1; }
  if (!cls || !split_rank || !xyz || !raw_scaling || !raw_quat || !raw_opacity || !src_idx || !kind || !scratch_alloc ||
      (n_split > 0 && !samples)) { set_error("densify_plan: null pointer"); return 1; }
  const size_t n4 = (size_t)P * 4;
  char* sc = scratch_alloc(scratch_user, 2 * align_up(n4 * 4) + scan_scratch_bytes(n4) + 512);
  if (!sc) { set_error("densify_plan: scratch allocation failed"); return 1; }
  Carver cv(sc);
  uint32_t* keep = cv.take<uint32_t>(n4);
  uint32_t* dest = cv.take<uint32_t>(n4);
  uint32_t* total = cv.take<uint32_t>(1);
  char* scan_scratch = cv.take<char>(scan_scratch_bytes(n4));
  densify_plan_kernel<<<grid1(P), 256, 0, s>>>(P, cls, split_rank, n_split, xyz, raw_scaling, raw_quat, raw_opacity,
                                               samples, min_opacity, big_ws, has_zmin, zmin, num_replica, keep);
  if (exclusive_scan_u32(keep, nullptr, dest, total, n4, scan_scratch, false, s)) return 1;
  densify_scatter_kernel<<<grid1(4 * P), 256, 0, s>>>(P, keep, dest, src_idx, kind);
  uint32_t h = 0;
  cudaError_t e = cudaMemcpyAsync(&h, total, 4, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);      // the caller sizes the new model
  if (e != cudaSuccess) { set_error("[CUDA ERROR] densify_plan: %s", cudaGetErrorString(e)); return 1; }
  *n_out = (int64_t)h;
  return check_launch("densify_plan", false, s, 5);
}

extern "C" int dgr_gather_rows(int64_t n_out, int32_t row_words, const void* src, void* dst, const int32_t* src_idx,
                               const uint8_t* kind, int32_t zero_new, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (n_out <= 0 || row_words <= 0) return 0;
  if (!src || !dst || !src_idx) { set_error("gather_rows: null pointer"); return 1; }
  const unsigned long long total = (unsigned long long)n_out * (unsigned long long)row_words;
  unsigned long long g = (total + 255) / 256;
  if (g > 148ull * 32) g = 148ull * 32;
  gather_rows_kernel<<<(unsigned)g, 256, 0, s>>>((const uint32_t*)src, (uint32_t*)dst, (uint32_t)row_words, src_idx, kind,
                                                 zero_new, total);
  return check_launch("gather_rows", false, s);
}

extern "C" int dgr_densify_split_fixup(int64_t n_out, const int32_t* src_idx, const uint8_t* kind,
                                       const uint32_t* split_rank, int64_t n_split, const float* xyz,
                                       const float* raw_scaling, const float* raw_quat, const float* samples,
                                       int32_t num_replica, float* xyz_out, float* scaling_out, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (n_out <= 0 || n_split <= 0) return 0;
  if (!src_idx || !kind || !split_rank || !xyz || !raw_scaling || !raw_quat || !samples || !xyz_out || !scaling_out) {
    set_error("densify_split_fixup: null pointer");
    return 1;
  }
  densify_split_fixup_kernel<<<grid1(n_out), 256, 0, s>>>(n_out, src_idx, kind, split_rank, n_split, xyz, raw_scaling,
                                                          raw_quat, samples, num_replica, xyz_out, scaling_out);
  return check_launch("densify_split_fixup", false, s);
}

extern "C" int dgr_reset_opacity(int64_t P, float* raw_opacity, float* exp_avg, float* exp_avg_sq, float cap, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (P <= 0) return 0;
  if (!raw_opacity) { set_error("reset_opacity: null pointer"); return 1; }
  reset_opacity_kernel<<<grid1(P), 256, 0, s>>>(P, raw_opacity, exp_avg, exp_avg_sq, cap);
  return check_launch("reset_opacity", false, s);
}

extern "C" int dgr_densification_stats(int64_t P, const float* grad_means2D, const int32_t* radii, float* grad_accum,
                                       float* denom, float* max_radii2D, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (P <= 0) return 0;
  if (!grad_means2D || !radii || !grad_accum || !denom) { set_error("densification_stats: null pointer"); return 1; }
  densify_stats_kernel<<<grid1(P), 256, 0, s>>>(P, grad_means2D, radii, grad_accum, denom, max_radii2D);
  return check_launch("densification_stats", false, s);
}
