// Shared helpers for the sm_100a DOGS hot-path kernels.  No torch headers here:
// every translation unit compiles in seconds with plain nvcc.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/dogs_b200.h"

#define TILE_X 16   // cuda_rasterizer/config.h:16-17
#define TILE_Y 16
#define TILE_PIX (TILE_X * TILE_Y)

namespace dogs {

void set_error(const char* fmt, ...);
int check_launch(const char* what, bool debug, cudaStream_t s, int kernels = 1);

static inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// Bump carving of one caller-allocated buffer.
struct Carver {
  char* base; size_t off;
  explicit Carver(char* b) : base(b), off(0) {}
  template <typename T> T* take(size_t n) {
    T* p = reinterpret_cast<T*>(base + off);
    off = align_up(off + n * sizeof(T));
    return p;
  }
};

// ---- 3x3 column-major matrix with the multiplication order of GLM 0.9.9.9 --------
// (m[c][r]; R[c][r] = A[0][r]*B[c][0] + A[1][r]*B[c][1] + A[2][r]*B[c][2],
//  glm/detail/type_mat3x3.inl:486-519).  The scalar data-flow is kept identical to
// the reference so that nvcc's FMA contraction produces the same bits.
struct M3 {
  float m[3][3];
};
__device__ __forceinline__ M3 m3(float a, float b, float c, float d, float e, float f,
                                 float g, float h, float i) {
  M3 r;
  r.m[0][0] = a; r.m[0][1] = b; r.m[0][2] = c;
  r.m[1][0] = d; r.m[1][1] = e; r.m[1][2] = f;
  r.m[2][0] = g; r.m[2][1] = h; r.m[2][2] = i;
  return r;
}
__device__ __forceinline__ M3 operator*(const M3& A, const M3& B) {
  M3 R;
#pragma unroll
  for (int c = 0; c < 3; c++)
#pragma unroll
    for (int r = 0; r < 3; r++)
      R.m[c][r] = A.m[0][r] * B.m[c][0] + A.m[1][r] * B.m[c][1] + A.m[2][r] * B.m[c][2];
  return R;
}
__device__ __forceinline__ M3 transpose(const M3& A) {
  M3 R;
#pragma unroll
  for (int c = 0; c < 3; c++)
#pragma unroll
    for (int r = 0; r < 3; r++) R.m[c][r] = A.m[r][c];
  return R;
}

struct V3 { float x, y, z; };
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ V3 operator*(float s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
__device__ __forceinline__ V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ V3 operator/(V3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
__device__ __forceinline__ V3& operator+=(V3& a, V3 b) { a.x += b.x; a.y += b.y; a.z += b.z; return a; }
__device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

// SH basis constants (cuda_rasterizer/auxiliary.h:21-38)
#define SH_C0 0.28209479177387814f
#define SH_C1 0.4886025119029199f
#define SH_C2_0 1.0925484305920792f
#define SH_C2_1 -1.0925484305920792f
#define SH_C2_2 0.31539156525252005f
#define SH_C2_3 -1.0925484305920792f
#define SH_C2_4 0.5462742152960396f
#define SH_C3_0 -0.5900435899266435f
#define SH_C3_1 2.890611442640554f
#define SH_C3_2 -0.4570457994644658f
#define SH_C3_3 0.3731763325901154f
#define SH_C3_4 -0.4570457994644658f
#define SH_C3_5 1.445305721320277f
#define SH_C3_6 -0.5900435899266435f

// auxiliary.h:40-43 — the literals are double, so this is evaluated in FP64.
__device__ __forceinline__ float ndc2Pix(float v, int S) { return ((v + 1.0) * S - 1.0) * 0.5; }

// auxiliary.h:69-88
__device__ __forceinline__ float3 xform4x3(const float3& p, const float* m) {
  return {m[0] * p.x + m[4] * p.y + m[8] * p.z + m[12],
          m[1] * p.x + m[5] * p.y + m[9] * p.z + m[13],
          m[2] * p.x + m[6] * p.y + m[10] * p.z + m[14]};
}
__device__ __forceinline__ float4 xform4x4(const float3& p, const float* m) {
  return {m[0] * p.x + m[4] * p.y + m[8] * p.z + m[12],
          m[1] * p.x + m[5] * p.y + m[9] * p.z + m[13],
          m[2] * p.x + m[6] * p.y + m[10] * p.z + m[14],
          m[3] * p.x + m[7] * p.y + m[11] * p.z + m[15]};
}

// auxiliary.h:45-55
__device__ __forceinline__ void get_rect(const float2 p, int max_radius, uint2& rmin, uint2& rmax,
                                         uint32_t gx, uint32_t gy) {
  rmin = {min(gx, (uint32_t)max((int)0, (int)((p.x - max_radius) / TILE_X))),
          min(gy, (uint32_t)max((int)0, (int)((p.y - max_radius) / TILE_Y)))};
  rmax = {min(gx, (uint32_t)max((int)0, (int)((p.x + max_radius + TILE_X - 1) / TILE_X))),
          min(gy, (uint32_t)max((int)0, (int)((p.y + max_radius + TILE_Y - 1) / TILE_Y)))};
}

// Activations the reference applies in torch before each render (gaussian_splat_model.py:148-153),
// for DgrParams::raw_inputs.  Written to reproduce the torch CUDA kernels: sigmoid = 1/(1+exp(-x))
// and exp with the full-precision expf and an IEEE divide; F.normalize(q, dim=1) = q / max(|q|, 1e-12)
// with |q| = sqrt of the uncontracted squares summed in torch's reduction order.
__device__ __forceinline__ float act_sigmoid(float x) { return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x))); }
__device__ __forceinline__ float3 act_exp3(float3 s) { return {expf(s.x), expf(s.y), expf(s.z)}; }
__device__ __forceinline__ float quat_norm(float4 q) {
  // torch 2.11's reduce kernel sums a 4-wide row as (x0^2 + x2^2) + (x1^2 + x3^2) (two shuffle steps);
  // measured with tools/gpu_act_compare.py, where the other orders differ in ~4 % of the rows
  const float ss = __fadd_rn(__fadd_rn(__fmul_rn(q.x, q.x), __fmul_rn(q.z, q.z)),
                             __fadd_rn(__fmul_rn(q.y, q.y), __fmul_rn(q.w, q.w)));
  return fmaxf(__fsqrt_rn(ss), 1e-12f);
}
__device__ __forceinline__ float4 act_normalize(float4 q) {
  const float n = quat_norm(q);
  return {__fdiv_rn(q.x, n), __fdiv_rn(q.y, n), __fdiv_rn(q.z, n), __fdiv_rn(q.w, n)};
}

// Per-tile reachability test of the key-duplication stage
// (rasterizer_impl.cu:52-100 max_contrib_power_rect_gaussian_float<15,15> + :151-173).
// Returns true iff the Gaussian can reach alpha >= 1/255 somewhere in tile (tx,ty).
// The two reciprocals depend on the Gaussian only; callers that walk a whole rect compute them once
// (tile_rcps) — same operations on the same operands, so the decisions are unchanged.
__device__ __forceinline__ float2 tile_rcps(const float4 co) {
  return {__frcp_rn((uint32_t)((TILE_X - 1) * (TILE_X - 1)) * co.x),
          __frcp_rn((uint32_t)((TILE_Y - 1) * (TILE_Y - 1)) * co.z)};
}
__device__ __forceinline__ bool tile_reachable(const float4 co, const float2 mean, int tx, int ty,
                                               float thr, const float2 rcps) {
  const float rminx = (float)(tx * TILE_X), rminy = (float)(ty * TILE_Y);
  const float rmaxx = (float)((tx + 1) * TILE_X - 1), rmaxy = (float)((ty + 1) * TILE_Y - 1);
  const float x_min_diff = rminx - mean.x;
  const float x_left = x_min_diff > 0.0f;
  const float not_in_x_range = x_left + (mean.x > rmaxx);
  const float y_min_diff = rminy - mean.y;
  const float y_above = y_min_diff > 0.0f;
  const float not_in_y_range = y_above + (mean.y > rmaxy);
  float max_contrib_power = 0.0f;
  if ((not_in_y_range + not_in_x_range) > 0.0f) {
    const float px = x_left * rminx + (1.0f - x_left) * rmaxx;
    const float py = y_above * rminy + (1.0f - y_above) * rmaxy;
    const float dx = copysignf(float(TILE_X - 1), x_min_diff);
    const float dy = copysignf(float(TILE_Y - 1), y_min_diff);
    const float diffx = mean.x - px;
    const float diffy = mean.y - py;
    const float rcp_dxdxcox = rcps.x;
    const float rcp_dydycoz = rcps.y;
    const float tx_ = not_in_y_range * __saturatef((dx * co.x * diffx + dx * co.y * diffy) * rcp_dxdxcox);
    const float ty_ = not_in_x_range * __saturatef((dy * co.y * diffx + dy * co.z * diffy) * rcp_dydycoz);
    const float mx = px + tx_ * dx, my = py + ty_ * dy;
    const float ddx = mean.x - mx, ddy = mean.y - my;
    max_contrib_power = 0.5f * (co.x * ddx * ddx + co.z * ddy * ddy) + co.y * ddx * ddy;
  }
  return max_contrib_power <= thr;
}
__device__ __forceinline__ bool tile_reachable(const float4 co, const float2 mean, int tx, int ty,
                                               float thr) {
  return tile_reachable(co, mean, tx, ty, thr, tile_rcps(co));
}

// Internal per-Gaussian record written by preprocess and gathered by the blend kernels.
// a = {conic.x, conic.y, conic.z, opacity}, b = {x, y, 1/depth, r}, c = {g, b, depth, ln(255*opacity)}
struct __align__(16) SplatRecord { float4 a, b, c; };

// The opt-in dynamic shared-memory limit is a per-device function attribute: set it once per device
// the library is used on (a process may drive several GPUs).
template <typename K> static void allow_big_smem(K kernel, bool (&done)[64]) {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !done[dev]) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    if (dev >= 0 && dev < 64) done[dev] = true;
  }
}

// ---- bulk asynchronous copies (TMA engine, 1-D: cp.async.bulk; SASS UBLKCP) + mbarrier ------------------------
// One elected thread moves a contiguous block between global and shared memory without LSU instructions or
// registers; completion of a load is signalled on an mbarrier (complete_tx byte count), so the rest of the CTA
// computes while the block is in flight.  Addresses and sizes must be multiples of 16 bytes.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t arrivals) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// arm the barrier with the byte count of the copy and start it
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "MBAR_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@!p bra MBAR_WAIT_%=;\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// generic-proxy writes to shared memory -> visible to the async proxy (call in the writing threads, then barrier)
__device__ __forceinline__ void fence_smem_to_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// shared -> global; returns when the shared source may be reused (and the CTA may exit)
__device__ __forceinline__ void bulk_store_and_wait(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// ---- programmatic dependent launch ----------------------------------------------------------------------------
// The rasterizer is a chain of ~30 short kernels per step, each depending on the previous one.  Launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, a kernel may be scheduled while its predecessor is still
// draining; it calls pdl_wait() (griddepcontrol.wait: returns once the predecessor grid has completed and its
// writes are visible) before it touches anything the predecessor produced, so the launch latency and the
// prologue overlap the predecessor's tail.  pdl_trigger() lets the successor be scheduled as early as possible.
// Both are no-ops for a kernel launched the ordinary way.  DOGS_B200_PDL=0 switches the attribute off.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool pdl_enabled();
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                                     Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---- stage launchers (defined in the .cu files) -----------------------------------
struct GeomState {
  SplatRecord* rec; uint8_t* clamped; uint32_t* tiles; uint32_t* depth_key;
  uint32_t* sorted_idx;    // Gaussian ids in depth order
  uint32_t* offsets;       // exclusive scan of tiles in depth order
  // per Gaussian: {rect word, instance count, reachable-tile mask (bit = row-major index in the rect)};
  // rect word = min.x | min.y << 11 | width << 22, width 0 = rect larger than 64 tiles (mask unused)
  uint4* emit_rec;
  uint4* sorted_emit;      // the same in depth order, .y replaced by the Gaussian id
  uint32_t* sorted_cnt;    // instance counts in depth order (scan input)
  // device-side control words: [0] smallest, [1] largest depth key of the Gaussians that emit instances,
  // [2] number of tile instances R (exclusive-scan total), [3] 1 = R exceeded the binning capacity
  uint32_t* ctl;
  static size_t bytes(size_t P);
  static GeomState carve(char* p, size_t P);
};
struct BinState {
  uint32_t* point_list; uint32_t* tile_keys;
  // Sub-tile reachability found by the forward blend, reused by the backward: for every group of 32 consecutive
  // list entries of a tile, one word per 8x4 pixel block of the tile (8 words = 32 bytes per group; bit = entry
  // can reach the block, blend_common.cuh).  Group q of tile t sits in slot range.x/32 + t + q — the "+ t" keeps
  // the slots of neighbouring tiles apart although their lists do not start on multiples of 32.
  uint32_t* hits;
  static size_t hit_slots(size_t R, size_t T) { return R / 32 + T + 16; }
  static size_t bytes(size_t R, size_t T);
  static BinState carve(char* p, size_t R, size_t T);
};
struct ImgState {
  uint2* ranges; uint32_t* tile_order;   // tile ids, longest list first: the launch order of the blend kernels
  float* final_T; uint32_t* n_contrib;
  float* pix_color; float* pix_invdepth;  // forward outputs kept for the backward replay
  static size_t bytes(size_t N, size_t T);
  static ImgState carve(char* p, size_t N, size_t T);
};

// radix_sort.cu
// Optional device-side control of a sort (passed to the kernels by value):
//   n_dev  element count in device memory, or null: the launch is sized for the host-side capacity `n`, the
//          kernels use min(*n_dev, n).
struct SortCtl { const uint32_t* n_dev; };
// number of low bits in which two keys can differ, given the smallest and largest of them (>= 1)
static inline int significant_key_bits(uint32_t kmin, uint32_t kmax) {
  if (kmin >= kmax) return 1;
  uint32_t d = kmin ^ kmax;
  int hb = 0;
  while (d) { hb++; d >>= 1; }
  return hb;
}
// max_passes: digit passes (8 bits each) the caller will ask for at most, 1..4
size_t sort_pairs_scratch_bytes(size_t n, int max_passes = 4);
// Stable LSD sort of (key,val) u32 pairs on bits [begin_bit,end_bit), 8 bits per pass, ping-pong
// between the (a) and (b) buffers starting from (a); *result_in_b tells where the result is.
int sort_pairs_u32(uint32_t* keys_a, uint32_t* vals_a, uint32_t* keys_b, uint32_t* vals_b,
                   size_t n, int begin_bit, int end_bit, char* scratch, bool debug, cudaStream_t s,
                   int* result_in_b, bool vals_are_iota = false,   // true: vals_a is treated as 0..n-1 and never read
                   const SortCtl* ctl = nullptr);
size_t scan_scratch_bytes(size_t n);
// out[i] = sum_{j<i} in[idx ? idx[j] : j]; total (may be null) receives the grand total.
int launch_block_base_scan(uint32_t* sums, uint32_t nb, uint32_t* total_out, bool debug, cudaStream_t s);
int exclusive_scan_u32(const uint32_t* in, const uint32_t* idx, uint32_t* out, uint32_t* total,
                       size_t n, char* scratch, bool debug, cudaStream_t s);


// rasterizer_fwd.cu
int launch_preprocess_fwd(const DgrParams& prm, const float* means3D, const float* dc,
                          const float* sh, const float* colors_precomp, const float* opacities,
                          const float* scales, const float* rotations, const float* cov3D_precomp,
                          int32_t* radii, GeomState& g, cudaStream_t s);
int launch_depth_keys(const DgrParams& prm, const GeomState& g, cudaStream_t s);
size_t gather_scan_scratch_bytes(size_t P);
int launch_gather_scan(const DgrParams& prm, const GeomState& g, const uint32_t* sorted_idx, char* scratch,
                       uint32_t* total_out, cudaStream_t s);
int launch_emit(const DgrParams& prm, const GeomState& g, const int32_t* radii, uint32_t* tile_keys,
                uint32_t* vals, uint32_t capacity, cudaStream_t s);
// R_cap: rows the launch is sized for; R_dev (nullable): actual count in device memory
int launch_tile_ranges(const DgrParams& prm, uint32_t R_cap, const uint32_t* R_dev, const uint32_t* tile_keys,
                       uint2* ranges, cudaStream_t s);
int launch_tile_order(const DgrParams& prm, const ImgState& im, cudaStream_t s);
int launch_blend_fwd(const DgrParams& prm, const GeomState& g, const BinState& b, const ImgState& im,
                     float* out_color, float* out_invdepth, cudaStream_t s);
int launch_activations(int P, const float* raw_op, const float* raw_sc, const float* raw_rot, float* op,
                       float* sc, float* rot, cudaStream_t s);
int launch_filter(const DgrParams& prm, const float* means3D, const float* scales,
                  const float* rotations, const float* cov3D_precomp, int32_t* radii, cudaStream_t s);
int launch_mark_visible(int P, const float* means3D, const float* view, uint8_t* present,
                        cudaStream_t s);
// rasterizer_bwd.cu
size_t bwd_accum_bytes(size_t P);
int launch_blend_bwd(const DgrParams& prm, const GeomState& g, const BinState& b, const ImgState& im,
                     const float* dL_dcolor, const float* dL_dinvdepth, float* accum,
                     cudaStream_t s);
int launch_preprocess_bwd(const DgrParams& prm, const float* means3D, const float* dc,
                          const float* sh, const float* opacities, const float* scales,
                          const float* rotations, const float* cov3D_precomp, const int32_t* radii,
                          const GeomState& g, const float* accum, float* dL_dmeans3D,
                          float* dL_dmeans2D, float* dL_ddc, float* dL_dsh,
                          float* dL_dcolors_precomp, float* dL_dopacity, float* dL_dscales,
                          float* dL_drotations, float* dL_dcov3D, double* pen_partial,
                          cudaStream_t s);
size_t bwd_penalty_partial_bytes(size_t P);   // per-CTA partial penalty sums (DgrAdmmPenalty::sums)

}  // namespace dogs
