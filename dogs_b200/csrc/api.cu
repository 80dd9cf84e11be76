// C-ABI entry points of the rasterizer (include/dogs_b200.h).  Host orchestration only:
// carves the caller-allocated buffers, launches the stage kernels on the caller's stream.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include "common.cuh"

namespace dogs {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

bool pdl_enabled() {
  static const bool on = !(getenv("DOGS_B200_PDL") && getenv("DOGS_B200_PDL")[0] == '0');
  return on;
}

static unsigned long long g_launches = 0;
unsigned long long launch_count() { return g_launches; }

// Optional per-stage device timing (dogs_stage_timing / dogs_stage_times): off by default; when on, the
// instrumented entry points bracket their stages with CUDA events on the caller's stream and synchronise.
static bool g_stage_timing = false;
static float g_stage_ms[16] = {0.f};
static int g_stage_n = 0;
static float g_fwd_stage_ms[16] = {0.f};
static int g_fwd_stage_n = 0;

// records an event per forward stage boundary when stage timing is on
struct StageClock {
  cudaEvent_t ev[12];
  int n = 0;
  bool on;
  cudaStream_t s;
  StageClock(bool enabled, cudaStream_t stream) : on(enabled), s(stream) { mark(); }
  void mark() {
    if (!on || n >= 12) return;
    cudaEventCreate(&ev[n]);
    cudaEventRecord(ev[n], s);
    n++;
  }
  void finish(float* out, int* count) {
    if (!on) return;
    cudaEventSynchronize(ev[n - 1]);
    for (int i = 0; i + 1 < n; i++) cudaEventElapsedTime(&out[i], ev[i], ev[i + 1]);
    *count = n - 1;
    for (int i = 0; i < n; i++) cudaEventDestroy(ev[i]);
  }
};

int check_launch(const char* what, bool debug, cudaStream_t s, int kernels) {
  g_launches += (unsigned long long)kernels;
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess && debug) e = cudaStreamSynchronize(s);  // CHECK_CUDA, auxiliary.h:177-184
  if (e != cudaSuccess) {
    set_error("[CUDA ERROR] in %s: %s", what, cudaGetErrorString(e));
    return 1;
  }
  return 0;
}

// ---- buffer layouts ---------------------------------------------------------------------
size_t GeomState::bytes(size_t P) {
  Carver c(nullptr);
  c.take<SplatRecord>(P); c.take<uint8_t>(P); c.take<uint32_t>(P); c.take<uint32_t>(P);
  c.take<uint32_t>(P); c.take<uint32_t>(P); c.take<uint4>(P); c.take<uint4>(P); c.take<uint32_t>(P);
  c.take<uint32_t>(16);
  return c.off;
}
GeomState GeomState::carve(char* p, size_t P) {
  Carver c(p);
  GeomState g;
  g.rec = c.take<SplatRecord>(P);
  g.clamped = c.take<uint8_t>(P);
  g.tiles = c.take<uint32_t>(P);
  g.depth_key = c.take<uint32_t>(P);
  g.sorted_idx = c.take<uint32_t>(P);
  g.offsets = c.take<uint32_t>(P);
  g.emit_rec = c.take<uint4>(P);
  g.sorted_emit = c.take<uint4>(P);
  g.sorted_cnt = c.take<uint32_t>(P);
  g.ctl = c.take<uint32_t>(16);
  return g;
}
size_t BinState::bytes(size_t R, size_t T) {
  Carver c(nullptr);
  c.take<uint32_t>(R); c.take<uint32_t>(R); c.take<uint32_t>(8 * hit_slots(R, T));
  return c.off;
}
BinState BinState::carve(char* p, size_t R, size_t T) {
  Carver c(p);
  BinState b;
  b.point_list = c.take<uint32_t>(R);
  b.tile_keys = c.take<uint32_t>(R);
  b.hits = c.take<uint32_t>(8 * hit_slots(R, T));
  return b;
}
size_t ImgState::bytes(size_t N, size_t T) {
  Carver c(nullptr);
  c.take<uint2>(T); c.take<uint32_t>(T); c.take<float>(N); c.take<uint32_t>(N); c.take<float>(3 * N); c.take<float>(N);
  return c.off;
}
ImgState ImgState::carve(char* p, size_t N, size_t T) {
  Carver c(p);
  ImgState im;
  im.ranges = c.take<uint2>(T);
  im.tile_order = c.take<uint32_t>(T);
  im.final_T = c.take<float>(N);
  im.n_contrib = c.take<uint32_t>(N);
  im.pix_color = c.take<float>(3 * N);
  im.pix_invdepth = c.take<float>(N);
  return im;
}

// Pinned host word + side stream + events used to read the instance count back without stalling the stream
// the kernels run on.  One per (host thread, device), created on first use.
struct HostReadback { uint32_t* pinned; cudaStream_t side; cudaEvent_t ev_ready, ev_done; bool ok; };
static HostReadback* readback() {
  static thread_local HostReadback slots[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  HostReadback& r = slots[dev];
  if (!r.ok) {
    if (cudaHostAlloc((void**)&r.pinned, 64, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    if (cudaStreamCreateWithFlags(&r.side, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&r.ev_ready, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&r.ev_done, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    r.ok = true;
  }
  return &r;
}

// getHigherMsb (rasterizer_impl.cu:35-50): number of bits needed for tile ids, 2^bit > T.
static int higher_msb(uint32_t n) {
  int bit = 0;
  while (bit < 32 && (n >> bit)) bit++;
  return bit;
}

static int validate(const DgrParams* prm) {
  if (!prm) { set_error("DgrParams is null"); return 1; }
  if (prm->P < 0 || prm->W <= 0 || prm->H <= 0) {
    set_error("invalid sizes P=%d W=%d H=%d", prm->P, prm->W, prm->H);
    return 1;
  }
  if (prm->D < 0 || prm->D > 3) { set_error("sh_degree %d out of range 0..3", prm->D); return 1; }
  if (!prm->bg || !prm->viewmatrix || !prm->projmatrix || !prm->campos) {
    set_error("bg/viewmatrix/projmatrix/campos must be device pointers");
    return 1;
  }
  return 0;
}

}  // namespace dogs

using namespace dogs;

extern "C" {

const char* dogs_last_error(void) { return g_err; }
int dogs_abi_version(void) { return DOGS_B200_ABI_VERSION; }
unsigned long long dogs_launch_count(void) { return dogs::launch_count(); }

int dgr_layout(int32_t P, int32_t W, int32_t H, int32_t R, DgrLayout* out) {
  if (!out) return 1;
  const size_t N = (size_t)W * H;
  const size_t T = (size_t)((W + TILE_X - 1) / TILE_X) * ((H + TILE_Y - 1) / TILE_Y);
  GeomState g = GeomState::carve(nullptr, P);
  BinState b = BinState::carve(nullptr, R, T);
  ImgState im = ImgState::carve(nullptr, N, T);
  out->geom_record = (size_t)g.rec;
  out->geom_clamped = (size_t)g.clamped;
  out->geom_tiles = (size_t)g.tiles;
  out->geom_bytes = GeomState::bytes(P);
  out->bin_point_list = (size_t)b.point_list;
  out->bin_tile_keys = (size_t)b.tile_keys;
  out->bin_bytes = BinState::bytes(R, T);
  out->img_ranges = (size_t)im.ranges;
  out->img_final_T = (size_t)im.final_T;
  out->img_n_contrib = (size_t)im.n_contrib;
  out->img_bytes = ImgState::bytes(N, T);
  return 0;
}

int dgr_deferred_check(const uint32_t* words, int32_t instance_capacity, int32_t depth_key_bits,
                       int32_t* num_rendered, int32_t* key_bits) {
  if (!words) return 0;
  const int kb = significant_key_bits(words[0], words[1]);
  if (num_rendered) *num_rendered = (int32_t)words[2];
  if (key_bits) *key_bits = kb;
  const int assumed = (depth_key_bits >= 1 && depth_key_bits <= 32) ? depth_key_bits : 32;
  return (words[2] <= (uint32_t)instance_capacity && kb <= assumed) ? 1 : 0;
}

int dgr_forward(const DgrParams* prm, const float* means3D, const float* dc, const float* sh,
                const float* colors_precomp, const float* opacities, const float* scales,
                const float* rotations, const float* cov3D_precomp, float* out_color,
                float* out_invdepth, int32_t* radii, dogs_alloc_fn geom_alloc, void* geom_user,
                dogs_alloc_fn binning_alloc, void* binning_user, dogs_alloc_fn image_alloc,
                void* image_user, dogs_alloc_fn scratch_alloc, void* scratch_user,
                DgrForwardInfo* info, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (validate(prm)) return 1;
  const DgrParams& p = *prm;
  if (!means3D && p.P > 0) { set_error("means3D is null"); return 1; }
  if ((colors_precomp == nullptr) == (dc == nullptr && sh == nullptr) && p.P > 0) {
    set_error("Please provide excatly one of either SHs or precomputed colors!");
    return 1;
  }
  if (colors_precomp == nullptr && p.P > 0 && (dc == nullptr || (p.D > 0 && (sh == nullptr || p.M < (p.D + 1) * (p.D + 1) - 1)))) {
    set_error("SH input incomplete: need dc[P,1,3] and sh[P,M>=%d,3] for degree %d", (p.D + 1) * (p.D + 1) - 1, p.D);
    return 1;
  }
  if (((scales == nullptr || rotations == nullptr) == (cov3D_precomp == nullptr)) && p.P > 0) {
    set_error("Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!");
    return 1;
  }
  if (!geom_alloc || !binning_alloc || !image_alloc || !scratch_alloc) {
    set_error("allocation callbacks must not be null");
    return 1;
  }
  const size_t P = p.P, N = (size_t)p.W * p.H;
  const uint32_t gx = (p.W + TILE_X - 1) / TILE_X, gy = (p.H + TILE_Y - 1) / TILE_Y;
  const size_t T = (size_t)gx * gy;
  const bool dbg = p.debug != 0;

  const size_t geom_bytes = GeomState::bytes(P), img_bytes = ImgState::bytes(N, T);
  char* geom_p = geom_alloc(geom_user, geom_bytes);
  char* img_p = image_alloc(image_user, img_bytes);
  if (!geom_p || !img_p) { set_error("allocation callback returned null"); return 1; }
  GeomState g = GeomState::carve(geom_p, P);
  ImgState im = ImgState::carve(img_p, N, T);

  // The instance count R is produced on the device (exclusive-scan total).  Sizing the binning buffers by it
  // costs a host round trip in the middle of the pass (the reference has two: rasterizer_impl.cu:423,469).  With
  // DgrParams::instance_capacity > 0 the buffers are allocated for that many rows BEFORE anything is launched,
  // every kernel after the scan takes R from device memory, the whole pass is enqueued without waiting, and R is
  // copied to pinned host memory on a side stream while the emit / tile-sort / blend kernels run.  The host
  // looks at it only at the end: the GPU never idles.  R > capacity (kernels clamp, nothing is written out of
  // bounds) re-runs the stages after the scan with exactly sized buffers.
  uint32_t R = 0, rows = 0;
  int key_bits = 0;
  size_t scratch1_bytes = 0, scratch2_bytes = 0;
  StageClock clk(g_stage_timing, s);
  // deferred mode (DgrParams::deferred_check): nothing below waits for the device or touches a second stream
  const bool deferred = p.deferred_check != nullptr;
  if (deferred && p.instance_capacity <= 0) { set_error("forward: deferred_check needs instance_capacity > 0"); return 1; }
  HostReadback* rb = deferred ? nullptr : readback();
  if (!deferred && !rb) { set_error("forward: cannot create the pinned read-back slot"); return 1; }
  const int bits = higher_msb((uint32_t)T);
  const int passes = (bits + 7) / 8;
  BinState b = BinState::carve(nullptr, 0, T);
  size_t bin_bytes = 0;

  // stages after the scan, for a binning buffer of `cap` rows; dev_count: take R from g.ctl[2]
  auto bin_and_blend = [&](uint32_t cap, bool dev_count) -> int {
    bin_bytes = BinState::bytes(cap, T);
    char* bin_p = binning_alloc(binning_user, bin_bytes > 0 ? bin_bytes : 256);
    if (!bin_p) { set_error("binning allocation failed"); return 1; }
    b = BinState::carve(bin_p, cap, T);
    rows = cap;
    const uint32_t* R_dev = dev_count ? g.ctl + 2 : nullptr;
    if (cap > 0) {
      scratch2_bytes = align_up((size_t)cap * 4) * 2 + sort_pairs_scratch_bytes(cap, passes) + 256;
      char* sc = scratch_alloc(scratch_user, scratch2_bytes);
      if (!sc) { set_error("scratch allocation failed"); return 1; }
      Carver cv(sc);
      uint32_t* keys_u = cv.take<uint32_t>(cap);
      uint32_t* vals_u = cv.take<uint32_t>(cap);
      char* sort_scratch = cv.take<char>(sort_pairs_scratch_bytes(cap, passes));
      // arrange the ping-pong so that the last pass lands in the persistent binning buffers
      uint32_t *ka, *va, *kb, *vb;
      if (passes % 2 == 0) { ka = b.tile_keys; va = b.point_list; kb = keys_u; vb = vals_u; }
      else { ka = keys_u; va = vals_u; kb = b.tile_keys; vb = b.point_list; }
      if (launch_emit(p, g, radii, ka, va, cap, s)) return 1;
      clk.mark();   // [3] emit
      int in_b = 0;
      const SortCtl ctl_t{R_dev};
      if (sort_pairs_u32(ka, va, kb, vb, cap, 0, bits, sort_scratch, dbg, s, &in_b, false, &ctl_t)) return 1;
      clk.mark();   // [4] tile sort
    }
    if (launch_tile_ranges(p, cap, R_dev, b.tile_keys, im.ranges, s)) return 1;
    if (launch_tile_order(p, im, s)) return 1;
    clk.mark();     // [5] ranges + order
    if (launch_blend_fwd(p, g, b, im, out_color, out_invdepth, s)) return 1;
    clk.mark();     // [6] blend
    return 0;
  };

  if (P > 0) {
    cudaMemsetAsync(g.ctl, 0, 16 * sizeof(uint32_t), s);
    cudaMemsetAsync(g.ctl, 0xff, sizeof(uint32_t), s);          // ctl[0] = smallest depth key so far
    if (launch_preprocess_fwd(p, means3D, dc, sh, colors_precomp, opacities, scales, rotations,
                              cov3D_precomp, radii, g, s)) return 1;
    clk.mark();   // [0] preprocess
    // depth order of the Gaussians (ties broken by index: the sort is stable)
    scratch1_bytes = align_up(P * 4) * 2 + sort_pairs_scratch_bytes(P) + gather_scan_scratch_bytes(P) + 256;
    char* sc = scratch_alloc(scratch_user, scratch1_bytes);
    if (!sc) { set_error("scratch allocation failed"); return 1; }
    Carver cv(sc);
    uint32_t* keys_b = cv.take<uint32_t>(P);
    uint32_t* vals_b = cv.take<uint32_t>(P);
    char* sort_scratch = cv.take<char>(sort_pairs_scratch_bytes(P));
    char* scan_scratch = cv.take<char>(gather_scan_scratch_bytes(P));
    // Depth order.  Positive floats of one scene share their top bits, so only the low `key_bits` bits — up to
    // the highest bit in which the smallest and the largest key differ — need sorting (3 passes of 8 bits instead
    // of 4 for depths within a factor of ~2-4).  The host cannot know that range yet, so it sorts on the bits the
    // CALLER expects (DgrParams::depth_key_bits, learnt from earlier frames), reads the true range back together
    // with the instance count, and repeats the depth stage on all 32 bits in the (rare) case it was wider.
    // Culled Gaussians carry the key 0xffffffff: on any bit range they sort last among equals and own no instances.
    auto depth_stage = [&](int key_bits) -> int {
      int in_b = 0;
      if (sort_pairs_u32(g.depth_key, g.sorted_idx, keys_b, vals_b, P, 0, key_bits, sort_scratch, dbg, s, &in_b, true)) return 1;
      clk.mark();   // [1] depth sort
      // one 16-byte gather per Gaussian brings everything the emit stage needs into depth order, and the same
      // kernel scans the instance counts (offsets, R): the emit kernel then streams sequentially
      return launch_gather_scan(p, g, in_b ? vals_b : g.sorted_idx, scan_scratch, g.ctl + 2, s);
    };
    const int bits_assumed = (p.depth_key_bits >= 1 && p.depth_key_bits <= 32) ? p.depth_key_bits : 32;
    if (depth_stage(bits_assumed)) return 1;
    if (deferred) {
      clk.mark();   // [2] gather + scan
      if (bin_and_blend((uint32_t)p.instance_capacity, true)) return 1;
      if (cudaMemcpyAsync(p.deferred_check, g.ctl, 16, cudaMemcpyDeviceToHost, s) != cudaSuccess) {
        set_error("[CUDA ERROR] deferred_check copy: %s", cudaGetErrorString(cudaGetLastError()));
        return 1;
      }
      clk.finish(g_fwd_stage_ms, &g_fwd_stage_n);
      if (info) {
        info->num_rendered = -1;
        info->num_visible = -1;
        info->binning_rows = (int32_t)rows;
        info->depth_key_bits = bits_assumed;
        info->geom_bytes = geom_bytes;
        info->binning_bytes = bin_bytes;
        info->image_bytes = img_bytes;
        info->scratch_bytes = scratch1_bytes > scratch2_bytes ? scratch1_bytes : scratch2_bytes;
      }
      return 0;
    }
    // {smallest key, largest key, R} -> pinned host memory on the side stream
    cudaError_t e = cudaEventRecord(rb->ev_ready, s);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(rb->side, rb->ev_ready, 0);
    if (e == cudaSuccess) e = cudaMemcpyAsync(rb->pinned, g.ctl, 12, cudaMemcpyDeviceToHost, rb->side);
    if (e == cudaSuccess) e = cudaEventRecord(rb->ev_done, rb->side);
    const bool optimistic = p.instance_capacity > 0;
    if (e == cudaSuccess && optimistic) {
      clk.mark();   // [2] gather + scan
      if (bin_and_blend((uint32_t)p.instance_capacity, true)) return 1;
    }
    if (e == cudaSuccess) e = cudaEventSynchronize(rb->ev_done);
    if (e != cudaSuccess) { set_error("[CUDA ERROR] reading num_rendered: %s", cudaGetErrorString(e)); return 1; }
    key_bits = significant_key_bits(rb->pinned[0], rb->pinned[1]);
    const bool redo_depth = key_bits > bits_assumed;
    if (redo_depth && (launch_depth_keys(p, g, s) || depth_stage(32))) return 1;
    R = rb->pinned[2];
    if (!optimistic) clk.mark();   // [2] gather + scan + host read of R
    if (!optimistic || redo_depth || R > (uint32_t)p.instance_capacity) {
      if (bin_and_blend(R, false)) return 1;
    }
  } else {
    if (bin_and_blend(0, false)) return 1;
  }
  clk.finish(g_fwd_stage_ms, &g_fwd_stage_n);

  if (info) {
    info->num_rendered = (int32_t)R;
    info->num_visible = -1;
    info->binning_rows = (int32_t)rows;
    info->depth_key_bits = key_bits;
    info->geom_bytes = geom_bytes;
    info->binning_bytes = bin_bytes;
    info->image_bytes = img_bytes;
    info->scratch_bytes = scratch1_bytes > scratch2_bytes ? scratch1_bytes : scratch2_bytes;
  }
  return 0;
}

int dgr_backward(const DgrParams* prm, int32_t num_rendered, const float* means3D, const float* dc,
                 const float* sh, const float* colors_precomp, const float* opacities,
                 const float* scales, const float* rotations, const float* cov3D_precomp,
                 const int32_t* radii, const char* geom, const char* binning, const char* image,
                 const float* dL_dcolor, const float* dL_dinvdepth_in, float* dL_dmeans3D,
                 float* dL_dmeans2D, float* dL_ddc, float* dL_dsh, float* dL_dcolors_precomp,
                 float* dL_dopacity, float* dL_dscales, float* dL_drotations, float* dL_dcov3D,
                 dogs_alloc_fn scratch_alloc, void* scratch_user, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (validate(prm)) return 1;
  const DgrParams& p = *prm;
  if (p.P == 0) return 0;
  float* stage_ms = g_stage_timing ? g_stage_ms : nullptr;
  if (!geom || !binning || !image || !dL_dcolor) { set_error("backward: null state or gradient"); return 1; }
  if (!dL_dmeans3D || !dL_dmeans2D || !dL_dopacity) { set_error("backward: null output"); return 1; }
  if (!scratch_alloc) { set_error("backward: null scratch callback"); return 1; }
  const size_t P = p.P, N = (size_t)p.W * p.H;
  const size_t T = (size_t)((p.W + TILE_X - 1) / TILE_X) * ((p.H + TILE_Y - 1) / TILE_Y);
  GeomState g = GeomState::carve(const_cast<char*>(geom), P);
  BinState b = BinState::carve(const_cast<char*>(binning), (size_t)num_rendered, T);
  ImgState im = ImgState::carve(const_cast<char*>(image), N, T);
  const bool pen_sums = p.admm && p.admm->zmu_xyz && p.admm->sums;
  float* accum = reinterpret_cast<float*>(
      scratch_alloc(scratch_user, bwd_accum_bytes(P) + (pen_sums ? bwd_penalty_partial_bytes(P) : 0)));
  if (!accum) { set_error("backward: scratch allocation failed"); return 1; }
  double* pen_partial = pen_sums ? reinterpret_cast<double*>(reinterpret_cast<char*>(accum) + bwd_accum_bytes(P)) : nullptr;
  cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
  if (stage_ms) {
    for (auto& e : ev) cudaEventCreate(&e);
    cudaEventRecord(ev[0], s);
  }
  if (launch_blend_bwd(p, g, b, im, dL_dcolor, dL_dinvdepth_in, accum, s)) return 1;
  if (stage_ms) cudaEventRecord(ev[1], s);
  const bool sh_path = colors_precomp == nullptr;
  int rc = launch_preprocess_bwd(p, means3D, dc, sh, opacities, scales, rotations, cov3D_precomp, radii,
                                 g, accum, dL_dmeans3D, dL_dmeans2D, sh_path ? dL_ddc : nullptr,
                                 sh_path ? dL_dsh : nullptr, dL_dcolors_precomp, dL_dopacity,
                                 cov3D_precomp ? nullptr : dL_dscales,
                                 cov3D_precomp ? nullptr : dL_drotations, dL_dcov3D, pen_partial, s);
  if (stage_ms) {
    cudaEventRecord(ev[2], s);
    cudaEventSynchronize(ev[2]);
    cudaEventElapsedTime(&stage_ms[0], ev[0], ev[1]);
    cudaEventElapsedTime(&stage_ms[1], ev[1], ev[2]);
    g_stage_n = 2;
    for (auto& e : ev) cudaEventDestroy(e);
  }
  return rc;
}

// Per-stage device timing for benchmarks (off by default).  While enabled, dgr_backward brackets its stages
// with CUDA events on the caller's stream and synchronises; dogs_stage_times returns the last call's stage
// times in ms: [0] accumulator clear + blend_bwd_kernel, [1] preprocess_bwd_kernel.
void dogs_stage_timing(int enable) { g_stage_timing = enable != 0; g_stage_n = 0; g_fwd_stage_n = 0; }
int dogs_stage_times(float* out_ms, int n) {
  if (!out_ms || n > g_stage_n) { set_error("dogs_stage_times: %d stage(s) recorded", g_stage_n); return 1; }
  for (int i = 0; i < n; i++) out_ms[i] = g_stage_ms[i];
  return 0;
}
// stages of the last dgr_forward: preprocess, depth sort, gather + scan (+ host read), emit, tile sort,
// ranges + order, blend; returns the number of stages copied (<= n), -1 when none was recorded
int dogs_forward_stage_times(float* out_ms, int n) {
  if (!out_ms || g_fwd_stage_n == 0) return -1;
  const int m = n < g_fwd_stage_n ? n : g_fwd_stage_n;
  for (int i = 0; i < m; i++) out_ms[i] = g_fwd_stage_ms[i];
  return m;
}

int dgr_apply_activations(int32_t P, const float* raw_opacities, const float* raw_scales,
                          const float* raw_rotations, float* opacities, float* scales, float* rotations,
                          void* stream) {
  if (P < 0 || (raw_opacities && !opacities) || (raw_scales && !scales) || (raw_rotations && !rotations)) {
    set_error("apply_activations: invalid arguments");
    return 1;
  }
  if (P == 0) return 0;
  return launch_activations(P, raw_opacities, raw_scales, raw_rotations, opacities, scales, rotations,
                            (cudaStream_t)stream);
}

int dgr_visible_filter(const DgrParams* prm, const float* means3D, const float* scales,
                       const float* rotations, const float* cov3D_precomp, int32_t* radii,
                       void* stream) {
  if (!prm || prm->P < 0 || prm->W <= 0 || prm->H <= 0 || !prm->viewmatrix || !prm->projmatrix) {
    set_error("visible_filter: invalid parameters");
    return 1;
  }
  if (prm->P == 0) return 0;
  if ((scales == nullptr || rotations == nullptr) == (cov3D_precomp == nullptr)) {
    set_error("Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!");
    return 1;
  }
  return launch_filter(*prm, means3D, scales, rotations, cov3D_precomp, radii, (cudaStream_t)stream);
}

int dgr_mark_visible(int32_t P, const float* means3D, const float* viewmatrix,
                     const float* projmatrix, uint8_t* present, void* stream) {
  (void)projmatrix;  // the reference's test only depends on the view-space depth
  if (P == 0) return 0;
  if (!means3D || !viewmatrix || !present) { set_error("mark_visible: null pointer"); return 1; }
  return launch_mark_visible(P, means3D, viewmatrix, present, (cudaStream_t)stream);
}

}  // extern "C"
