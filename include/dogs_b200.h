/*
 * dogs_b200.h — C-ABI of the B200-native (sm_100a) DOGS hot path.
 *
 * Plain C: raw device pointers, sizes, a cudaStream_t passed as void*.  No torch
 * types.  Every entry point returns 0 on success, non-zero on error;
 * dogs_last_error() returns the message of the calling thread's last failure
 * (the reference throws std::runtime_error / AT_ERROR instead:
 * submodules/diff-gaussian-rasterization/rasterize_points.cu:78-80,
 * cuda_rasterizer/rasterizer_impl.cu:382-385, cuda_rasterizer/auxiliary.h:177-184).
 *
 * All kernels are enqueued on the stream given by the caller (the reference uses
 * the legacy default stream everywhere).  The library owns no device memory:
 * persistent state and scratch come from caller-supplied allocation callbacks,
 * exactly like the reference's std::function<char*(size_t)> resize callbacks
 * (rasterize_points.cu:30-36,106-109; rasterizer_impl.cu:335-338).
 */
#ifndef DOGS_B200_H_INCLUDED
#define DOGS_B200_H_INCLUDED

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DOGS_B200_ABI_VERSION 3

/* Allocation callback: must return a device pointer to >= bytes bytes, aligned to
 * 256 B, valid until the caller decides to free it.  Mirrors resizeFunctional()
 * (rasterize_points.cu:30-36). */
typedef char* (*dogs_alloc_fn)(void* user, size_t bytes);

/* ---- rasterizer ----------------------------------------------------------------
 * Replaces CudaRasterizer::Rasterizer::{forward,backward,visible_filter,markVisible}
 * (cuda_rasterizer/rasterizer.h:22-113) and the glue in rasterize_points.cu:54-252. */
/* Optional fused ADMM penalty (conerf/trainers/slave_gaussian_trainer.py:161-202) for the consensus
 * groups that are direct rasterizer inputs: xyz = means3D, f_dc = dc, f_rest = sh always; scaling,
 * quaternion and opacity when the caller passes the consensus (raw) tensors themselves, i.e. with
 * DgrParams::raw_inputs.  The backward adds coef_g * (x_g - zmu_g) to the gradient of EVERY Gaussian
 * for each group whose zmu pointer is set, with zmu = z - u (admm_unpack_update) and
 * coef_g = rho_g / numel_g, in the pass that writes those gradients anyway.
 * sums (device double[6] in the order below, nullable) += sum (x_g - zmu_g)^2 for logging. */
typedef struct DgrAdmmPenalty {
  const float* zmu_xyz;    /* [P,3]   NULL disables the whole fusion */
  const float* zmu_dc;     /* [P,3]   nullable */
  const float* zmu_sh;     /* [P,M,3] nullable */
  const float* zmu_scale;  /* [P,3]   nullable; compared with the `scales` argument as passed */
  const float* zmu_rot;    /* [P,4]   nullable; compared with `rotations` as passed */
  const float* zmu_opacity;/* [P]     nullable; compared with `opacities` as passed */
  float coef_xyz, coef_dc, coef_sh, coef_scale, coef_rot, coef_opacity;
  double* sums;
} DgrAdmmPenalty;

typedef struct DgrParams {
  int32_t P;               /* number of Gaussians                                   */
  int32_t D;               /* active SH degree 0..3  (raster_settings.sh_degree)    */
  int32_t M;               /* SH "rest" coefficients per Gaussian = sh.size(1)      */
  int32_t W, H;            /* image size                                            */
  float tanfovx, tanfovy;
  float scale_modifier;
  float depth_threshold;   /* >0: grad_means2D *= min(1,(depth/thr)^2), __init__.py:171-185 */
  int32_t prefiltered;
  int32_t antialiasing;
  int32_t debug;           /* !=0: synchronise + check after every launch (CHECK_CUDA) */
  const float* bg;         /* [3]   device */
  const float* viewmatrix; /* [16]  device, column-major (camera.py:131)            */
  const float* projmatrix; /* [16]  device, column-major                            */
  const float* campos;     /* [3]   device                                          */
  int32_t ref_bucket_cull_quirk; /* backward only.  0 (default): exact gradients.  1: reproduce the
      reference's truncated backward: its per-tile max_contrib comes from cub::BlockReduce<uint32_t,256>
      instantiated 1-D on a 16x16 block (forward.cu:495-500), which only reduces the tile's first
      warp (pixel rows 0-1), so PerGaussianRenderCUDA (backward.cu:505-507) drops every 32-splat
      bucket starting at or after max(n_contrib of rows 0-1).  See DESIGN.md "Reference defect". */
  int32_t raw_inputs;      /* !=0: opacities / scales / rotations are the model's RAW parameters.  The
      activations the reference applies in torch before every render (gaussian_splat_model.py:148-153,
      179,202,259: sigmoid, exp, F.normalize) run on load inside preprocess / visible_filter, and
      dgr_backward returns gradients w.r.t. the raw tensors.  scales/rotations only
      when cov3D_precomp is null; opacity always. */
  const DgrAdmmPenalty* admm;    /* backward only, host pointer, nullable (see DgrAdmmPenalty) */
  /* forward only, nullable device pointers: LightGaussian importance statistics (count mode of the
   * reference's older rasterizer, old_diff-gaussian-rasterization/cuda_rasterizer/forward.cu:400-513,
   * consumed by conerf/model/gaussian_fields/prune.py:13-65): gaussians_count[P] = number of pixels
   * each Gaussian was blended into, important_score[P] = that count times its (anti-aliased) opacity. */
  int32_t* gaussians_count;
  float* important_score;
  /* forward only.  0: the binning buffers are sized by the instance count R, which the host reads in the middle
   * of the pass (one stream synchronisation; the reference has two).  > 0: capacity in tile instances — the
   * buffers are allocated for that many rows up front, the whole pass is enqueued at once with R taken from
   * device memory, and the host reads R only after everything has been launched, so the GPU never waits for it.
   * If R turns out larger (nothing is written out of bounds) the stages after the scan are re-run with exactly
   * sized buffers.  Results are identical either way.  A good value is ~1.25x the largest R seen so far. */
  int32_t instance_capacity;
  /* forward only.  Number of low bits of the depth keys (float bits of the view-space depth) the caller expects
   * to differ between the Gaussians of the frame, 1..32 (0 = 32): positive floats of one scene share their top
   * bits, and every 8 bits less is one pass of the depth sort less.  Take DgrForwardInfo::depth_key_bits of
   * earlier frames (their maximum).  The forward verifies it and repeats the depth stage on 32 bits if the
   * range was wider, so the results never depend on this value. */
  int32_t depth_key_bits;
  /* forward only, nullable: 4 words of PINNED host memory.  With instance_capacity > 0 and this pointer set the
   * forward does not wait for the device at all — no event, no second stream, so the call can be captured in a
   * CUDA graph: {smallest depth key, largest depth key, R, 0} are copied to the words in stream order and the
   * CALLER verifies, once the stream work has completed, that the assumptions held (dgr_deferred_check).  If they
   * did not, the outputs of that call (and of the backward that used its state) are invalid and the step must be
   * repeated with a larger capacity / more key bits.  DgrForwardInfo::num_rendered is -1 in this mode. */
  uint32_t* deferred_check;
} DgrParams;

/* Sizes / counters reported by dgr_forward (host side). */
typedef struct DgrForwardInfo {
  int32_t num_rendered;    /* tile instances that survive the per-tile cull (R_valid) */
  int32_t num_visible;     /* Gaussians with radii > 0 and at least one instance (not synced: -1) */
  size_t geom_bytes, binning_bytes, image_bytes, scratch_bytes;
  int32_t depth_key_bits;  /* bits in which the depth keys of this frame actually differed (see DgrParams) */
  int32_t binning_rows;    /* rows the binning buffer was carved for (= num_rendered, or the capacity that was used):
                              pass THIS as `num_rendered` to dgr_backward and dgr_layout */
} DgrForwardInfo;

/* Verdict on a deferred forward (DgrParams::deferred_check), to be called after the stream work has completed:
 * returns 1 if R <= instance_capacity and the depth keys differed in at most depth_key_bits bits (0 = 32), else 0.
 * num_rendered / key_bits (nullable) receive the measured values either way. */
int dgr_deferred_check(const uint32_t* words, int32_t instance_capacity, int32_t depth_key_bits,
                       int32_t* num_rendered, int32_t* key_bits);

/* Forward.  Nullable inputs follow the reference: exactly one of (dc+sh | colors_precomp),
 * exactly one of (scales+rotations | cov3D_precomp).
 *   out_color [3,H,W], out_invdepth [1,H,W], radii [P] are caller-allocated (need not be zeroed).
 *   geom/binning/image callbacks allocate the state kept for backward; scratch may be freed
 *   as soon as the call returns (stream-ordered).  Every buffer a callback hands out must stay valid until
 *   the call returns (the scratch callback is invoked more than once per call). */
int dgr_forward(const DgrParams* prm,
                const float* means3D, const float* dc, const float* sh,
                const float* colors_precomp, const float* opacities,
                const float* scales, const float* rotations, const float* cov3D_precomp,
                float* out_color, float* out_invdepth, int32_t* radii,
                dogs_alloc_fn geom_alloc, void* geom_user,
                dogs_alloc_fn binning_alloc, void* binning_user,
                dogs_alloc_fn image_alloc, void* image_user,
                dogs_alloc_fn scratch_alloc, void* scratch_user,
                DgrForwardInfo* info, void* stream);

/* Backward.  geom/binning/image are the buffers handed out during dgr_forward; num_rendered =
 * DgrForwardInfo::binning_rows of that forward.
 * Every gradient output is fully written (zeros where the reference relies on
 * torch::zeros, rasterize_points.cu:196-207), so the caller may pass torch.empty.
 * Nullable outputs: dL_ddc/dL_dsh (when colors_precomp was used), dL_dcolors_precomp,
 * dL_dscales/dL_drotations (when cov3D_precomp was used), dL_dcov3D, dL_dinvdepth_in. */
int dgr_backward(const DgrParams* prm, int32_t num_rendered,
                 const float* means3D, const float* dc, const float* sh,
                 const float* colors_precomp, const float* opacities,
                 const float* scales, const float* rotations, const float* cov3D_precomp,
                 const int32_t* radii,
                 const char* geom, const char* binning, const char* image,
                 const float* dL_dcolor, const float* dL_dinvdepth_in,
                 float* dL_dmeans3D, float* dL_dmeans2D, float* dL_ddc, float* dL_dsh,
                 float* dL_dcolors_precomp, float* dL_dopacity,
                 float* dL_dscales, float* dL_drotations, float* dL_dcov3D,
                 dogs_alloc_fn scratch_alloc, void* scratch_user, void* stream);

/* rasterize_gaussians_filter (rasterize_points.cu:274-329): radii only, no +0.3 low-pass. */
int dgr_visible_filter(const DgrParams* prm, const float* means3D, const float* scales,
                       const float* rotations, const float* cov3D_precomp,
                       int32_t* radii, void* stream);
/* mark_visible (rasterize_points.cu:254-272). present: [P] bytes (bool). */
int dgr_mark_visible(int32_t P, const float* means3D, const float* viewmatrix,
                     const float* projmatrix, uint8_t* present, void* stream);

/* The activations of DgrParams::raw_inputs as a stand-alone pass (gaussian_splat_model.py:148-153:
 * get_opacity = sigmoid(_opacity), get_scaling = exp(_scaling), get_quaternion = F.normalize(_quaternion)),
 * bit-identical to what preprocess applies and to the torch kernels; for callers that need the
 * activated values themselves (export, densification statistics).  Any of the three pairs may be null. */
int dgr_apply_activations(int32_t P, const float* raw_opacities, const float* raw_scales,
                          const float* raw_rotations, float* opacities, float* scales,
                          float* rotations, void* stream);

/* Introspection used by the parity tests (compare stage by stage with the reference's
 * opaque buffers).  Offsets in bytes inside the geom / binning / image buffers. */
typedef struct DgrLayout {
  size_t geom_record;      /* float4[3*P]: {cx,cy,cz,o} {x,y,1/depth,r} {g,b,depth,-}   */
  size_t geom_clamped;     /* uint8[P] bit c set => channel c clamped                  */
  size_t geom_tiles;       /* uint32[P] instances after cull                           */
  size_t geom_bytes;
  size_t bin_point_list;   /* uint32[R] Gaussian id per sorted instance                */
  size_t bin_tile_keys;    /* uint32[R] tile id per sorted instance                    */
  size_t bin_bytes;        /* (+ the forward's sub-tile reachability words read by the backward: ~1 byte per instance) */
  size_t img_ranges;       /* uint2[T]                                                 */
  size_t img_final_T;      /* float[N]                                                 */
  size_t img_n_contrib;    /* uint32[N]                                                */
  size_t img_bytes;
} DgrLayout;
int dgr_layout(int32_t P, int32_t W, int32_t H, int32_t num_rendered, DgrLayout* out);

/* SparseGaussianAdam.step -> adamUpdate (cuda_rasterizer/adam.cu:9-66). visible: [N] bytes. */
int dgr_sparse_adam(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                    const uint8_t* visible, float lr, float b1, float b2, float eps,
                    uint32_t N, uint32_t M, void* stream);

/* All parameter tensors of the model in one launch: replaces the loop of adamUpdate calls in
 * SparseGaussianAdam.step (__init__.py:303-332).  Visibility from a bool mask (visible, [N] bytes) or straight
 * from the rasterizer's radii (visible = radii > 0, gaussian_trainer.py:474-475); exactly one may be NULL.
 * Tensors must be 16-byte aligned; at most 8 per call.  Same arithmetic as dgr_sparse_adam (bit-identical). */
typedef struct DgrAdamTensor {
  float* param; const float* grad; float* exp_avg; float* exp_avg_sq;
  uint32_t M;              /* elements per Gaussian */
  float lr, eps;
} DgrAdamTensor;
int dgr_sparse_adam_multi(const DgrAdamTensor* tensors, int32_t n_tensors, const uint8_t* visible,
                          const int32_t* radii, float b1, float b2, uint32_t N, void* stream);

/* ---- stream compaction of the Gaussian SoA (prune / densify) ----------------
 * Replaces boolean-mask indexing of the 6 parameter tensors and their Adam state
 * (conerf/model/gaussian_fields/gaussian_splat_model.py:34-108,362-531). */
int dgr_compact_plan(int64_t P, const uint8_t* keep, uint32_t* offsets,
                     dogs_alloc_fn scratch_alloc, void* scratch_user, int64_t* count_out /* host */,
                     void* stream);
int dgr_compact_rows(int64_t P, int32_t row_words, const void* src, void* dst,
                     const uint8_t* keep, const uint32_t* offsets, void* stream);

/* ---- densification / pruning / opacity reset ---------------------------------
 * Replace GaussianSplatModel.densify_and_prune and the methods it calls, reset_opacity and
 * add_densification_stats (conerf/model/gaussian_fields/gaussian_splat_model.py:362-541) over the SoA.
 * All parameter tensors are the RAW model tensors (_scaling, _opacity, _quaternion).
 *
 * classify (:436-443, :462-470, :502-506): cls[i] bit 0 = clone, bit 1 = split with g = accum/denom (NaN -> 0),
 *   clone: |g| >= grad_threshold && max exp(scaling) <= scale_threshold; split: g >= grad_threshold && max > ...;
 *   split_rank[i] = number of split Gaussians before i; *n_split_out (host) their count (one stream sync: the
 *   caller draws the 2*n_split normal samples next, as the reference does with torch.normal).
 * plan (:472-527): decides which of the four possible output rows of every Gaussian survive the final prune
 *   (opacity < min_opacity; z < zmin when has_zmin; max activated scale > big_ws when big_ws >= 0) and numbers
 *   them in the reference's order [originals not split][clones][split replicas 0][split replicas 1];
 *   src_idx[j] / kind[j] (j < *n_out, arrays of 4*P entries) = source Gaussian / segment of output row j.
 * gather_rows: dst[j] = src[src_idx[j]] (rows of row_words 32-bit words), or 0 where zero_new && kind[j] != 0
 *   (Adam moments of appended rows, cat_tensors_to_optimizer :60-66).
 * split_fixup (:472-485): xyz and scaling of the split replicas: R(q/|q|) * sample + xyz, log(exp(s) / (0.8*n)). */
int dgr_densify_classify(int64_t P, const float* grad_accum, const float* denom, const float* raw_scaling,
                         float grad_threshold, float scale_threshold, uint8_t* cls, uint32_t* split_rank,
                         dogs_alloc_fn scratch_alloc, void* scratch_user, int64_t* n_split_out /* host */, void* stream);
int dgr_densify_plan(int64_t P, const uint8_t* cls, const uint32_t* split_rank, int64_t n_split,
                     const float* xyz, const float* raw_scaling, const float* raw_quat, const float* raw_opacity,
                     const float* samples /* [num_replica*n_split, 3] */, float min_opacity, float big_ws,
                     int32_t has_zmin, float zmin, int32_t num_replica, int32_t* src_idx, uint8_t* kind,
                     dogs_alloc_fn scratch_alloc, void* scratch_user, int64_t* n_out /* host */, void* stream);
int dgr_gather_rows(int64_t n_out, int32_t row_words, const void* src, void* dst, const int32_t* src_idx,
                    const uint8_t* kind /* nullable */, int32_t zero_new, void* stream);
int dgr_densify_split_fixup(int64_t n_out, const int32_t* src_idx, const uint8_t* kind, const uint32_t* split_rank,
                            int64_t n_split, const float* xyz, const float* raw_scaling, const float* raw_quat,
                            const float* samples, int32_t num_replica, float* xyz_out, float* scaling_out, void* stream);
/* reset_opacity (:362-367): opacity = inverse_sigmoid(min(sigmoid(opacity), cap)) in place; the tensor's Adam
 * moments (nullable) are zeroed (replace_tensor_to_optimizer :34-48). */
int dgr_reset_opacity(int64_t P, float* raw_opacity, float* exp_avg, float* exp_avg_sq, float cap, void* stream);
/* add_densification_stats (:529-541) + the trainer's max_radii2D update (gaussian_trainer.py:433-437) for the
 * Gaussians with radii > 0: accum += |grad_means2D.xy|, denom += 1, max_radii2D (nullable) = max(., radii). */
int dgr_densification_stats(int64_t P, const float* grad_means2D /* [P,3] */, const int32_t* radii,
                            float* grad_accum, float* denom, float* max_radii2D, void* stream);

/* ---- fused-SSIM (submodules/fused-ssim/ssim.cu:368-444) ------------------------- */
int fssim_forward(int32_t B, int32_t CH, int32_t H, int32_t W, float C1, float C2,
                  const float* img1, const float* img2, float* ssim_map,
                  float* dm_dmu1, float* dm_dsigma1_sq, float* dm_dsigma12, /* NULL when !train */
                  void* stream);
int fssim_backward(int32_t B, int32_t CH, int32_t H, int32_t W, float C1, float C2,
                   const float* img1, const float* img2, const float* dL_dmap,
                   const float* dm_dmu1, const float* dm_dsigma1_sq, const float* dm_dsigma12,
                   float* dL_dimg1, void* stream);

/* Fused photometric loss of the training step (conerf/trainers/gaussian_trainer.py:403-405):
 *   loss = (1 - lambda) * mean|img1 - img2| + lambda * (1 - mean SSIM(img1, img2)),
 * which the reference evaluates with ~20 torch elementwise/reduction launches around fused_ssim.
 * forward: ONE launch; writes the three derivative maps for the backward (NULL = inference) and the
 *   scalar loss_out (device).  workspace: 32 bytes of device memory, zeroed once by the caller; the
 *   kernel leaves it zeroed (one workspace per stream).
 * backward: ONE launch; dL_dimg1 = grad_loss * (w_ssim * dSSIMsum/dimg1 + w_l1 * sign(img1 - img2)) with
 *   w_ssim = -lambda/N, w_l1 = (1-lambda)/N, N = B*CH*H*W; grad_loss: device scalar, NULL = 1. */
int fssim_l1_loss_forward(int32_t B, int32_t CH, int32_t H, int32_t W, float C1, float C2,
                          float lambda_dssim, const float* img1, const float* img2,
                          float* dm_dmu1, float* dm_dsigma1_sq, float* dm_dsigma12,
                          void* workspace, float* loss_out, void* stream);
int fssim_l1_loss_backward(int32_t B, int32_t CH, int32_t H, int32_t W, float w_ssim, float w_l1,
                           const float* img1, const float* img2, const float* dm_dmu1,
                           const float* dm_dsigma1_sq, const float* dm_dsigma12,
                           const float* grad_loss, float* dL_dimg1, void* stream);

/* ---- simple-knn distCUDA2 (submodules/simple-knn/spatial.cu:17-35) --------------- */
int sknn_dist2(int32_t P, const float* points, float* mean_dist2,
               dogs_alloc_fn scratch_alloc, void* scratch_user, void* stream);

/* ---- ADMM consensus (conerf/trainers/{slave,master}_gaussian_trainer.py) --------
 * A Gaussian has 59 parameters in 6 groups (xyz 3, f_dc 3, f_rest 45, scaling 3,
 * quaternion 4, opacity 1); the 6 tensors stay separate (the model's own layout).  */
#define ADMM_GROUPS 6
typedef struct AdmmTensors { float* g[ADMM_GROUPS]; } AdmmTensors;
typedef struct AdmmConstTensors { const float* g[ADMM_GROUPS]; } AdmmConstTensors;

/* One consensus round of a block = admm_pack_rows -> exchange of the rows with the co-owning blocks
 * (NCCL send/recv between GPUs, nothing at all between blocks of the same GPU: the receiver reads the
 * sender's buffer) -> admm_consensus_update.  Replaces send_local_model / gaussian_splat_consensus /
 * broadcast_global_gaussian_splat / update_dual_variables / compute_primal_residual /
 * compute_dual_residual (master_gaussian_trainer.py:396-456,523-555; slave_gaussian_trainer.py:66-71,99-121)
 * for the Gaussians of one block.
 *
 * pack_rows: out[r*59 + :] = the 59 parameters of local Gaussian src_local[r], r < R (the rows for all
 * co-owners concatenated, each co-owner's rows ordered by global index). */
int admm_pack_rows(const AdmmConstTensors* x, const int32_t* src_local, int64_t R, float* out, void* stream);
/* The same gather with a destination per row: row r goes to dsts[dst_sel[r]] (<= 16 buffers of [rows, 59] floats)
 * at row dst_row[r].  With the co-owners' receive buffers mapped into this process (peer memory over NVLink /
 * NVSwitch, e.g. torch symmetric memory) the pack kernel writes every row straight where its consumer reads it:
 * pack and exchange are one kernel, and no NCCL call is made for the rows. */
int admm_pack_rows_to(const AdmmConstTensors* x, const int32_t* src_local, const uint8_t* dst_sel,
                      const int32_t* dst_row, int64_t R, float* const* dsts /* host array of device pointers */,
                      int32_t n_dsts, void* stream);
/* consensus_update over ALL n Gaussians of the block.
 *   shared_row[i] < 0: owned by this block only: z = x, u untouched (the reference's average with count 1).
 *   shared_row[i] = s: z = (sum of the addends of row s) / count[s]; u += relax * (x - z).
 *     addends[s*max_addends + j], j < max_addends, terminated by 0xffffffff: bits 31..28 = 15 -> this block's
 *     own x, else index into sources[]; bits 27..0 = row of that exchange buffer.  The addends are summed in
 *     the order listed; listing the owners in ascending block order reproduces the reference's
 *     plus_gaussians loop + average_gaussians (gaussian_splat_model.py:325-340) bit for bit.
 *   zmu (nullable) receives z - u for the per-iteration penalty.
 *   residuals (device double[12]) is ACCUMULATED: [0..5] sum (z-x)^2 per group (primal, master:396-432),
 *   [6..11] sum (z_prev-z)^2 / count (dual, master:435-456: every global Gaussian counted once over its owners).
 *   scratch: admm_consensus_scratch_bytes() of device memory, zeroed once by the caller (left zeroed). */
size_t admm_consensus_scratch_bytes(void);
int admm_consensus_update(const AdmmConstTensors* x, const AdmmTensors* u, const AdmmTensors* z,
                          const AdmmTensors* zmu /* nullable */, const int32_t* shared_row,
                          const uint32_t* addends, const uint8_t* count, int32_t max_addends, int64_t n,
                          const float* const* sources /* host array of device pointers */, int32_t n_sources,
                          float relax, void* scratch, double* residuals, void* stream);
/* penalty (slave:161-202): sums[g] += sum((x_g+u_g-z_g)^2) (device double[6]); u == NULL means that
 * `z` already holds z - u (the zmu tensors of admm_unpack_update: one pass less per iteration); if grad != NULL,
 * grad_g += coef[g]*(x_g+u_g-z_g) with coef[g] = rho_g/numel_g (host float[6]) — the gradient of
 * 0.5*rho_g*mse_loss(x_g+u_g, z_g). */
int admm_penalty(const AdmmConstTensors* x, const AdmmConstTensors* u, const AdmmConstTensors* z,
                 const AdmmTensors* grad, int64_t n, const float* coef /* host[6] */,
                 double* sums /* device[6] */, void* stream);

/* Block membership for the re-split after fusion (select_gaussians_in_each_block /
 * fuse_block_gaussians, conerf/trainers/master_gaussian_trainer.py:37-172, points_in_bbox2D
 * conerf/datasets/utils.py:186-205): bit b of mask[i] = Gaussian i is inside block b's closed 2-D box
 * boxes[b] = {min_x, min_y, max_x, max_y} (device double[n_blocks*4]) after the optional world->OBB
 * transform (host double[9], row-major 3x3 as passed to trimesh.transform_points; NULL = identity);
 * count[i] (nullable) = number of blocks holding i = torch.bincount of the reference.  <= 64 blocks. */
int admm_block_membership(int64_t n, const float* xyz, const double* world_to_obb,
                          int32_t n_blocks, const double* boxes, uint64_t* mask, int32_t* count,
                          void* stream);

const char* dogs_last_error(void);
int dogs_abi_version(void);
/* Number of kernels this library has launched so far in the process (bench.py reports the delta). */
unsigned long long dogs_launch_count(void);
/* Benchmark hook, off by default: while enabled, dgr_backward brackets its stages with CUDA events on the
 * caller's stream and synchronises; dogs_stage_times copies the last call's stage times (ms) to out_ms:
 * [0] accumulator clear + blend_bwd_kernel, [1] preprocess_bwd_kernel. */
void dogs_stage_timing(int enable);
int dogs_stage_times(float* out_ms, int n);
/* the same for the last dgr_forward: preprocess, depth sort, gather + scan, emit, tile sort, ranges + order,
 * blend (ms); returns the number of stages copied, -1 when none was recorded */
int dogs_forward_stage_times(float* out_ms, int n);

#ifdef __cplusplus
}
#endif
#endif
