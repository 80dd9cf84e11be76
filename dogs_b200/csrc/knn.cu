// simple-knn distCUDA2 (submodules/simple-knn/simple_knn.cu:185-221, spatial.cu:17-35): mean of the
// three smallest squared distances from every point to the others.  The search is exact, so the
// result depends only on the distance expression, which is kept as in the reference
// (d.x*d.x + d.y*d.y + d.z*d.z, simple_knn.cu:133-134) — outputs are bit-identical.
//
// Pipeline: bounding box (device-side, no host round trips; the reference does two blocking
// copies) -> 30-bit Morton codes -> radix sort (radix_sort.cu) -> points gathered into Morton
// order as float4 -> 1024-point box AABBs -> warp-cooperative pruned scan: the lanes of a warp are
// Morton neighbours, so the warp votes on each box and walks it together with broadcast loads
// instead of 32 divergent per-thread loops.
#include <float.h>
#include "common.cuh"

namespace dogs {

constexpr int KNN_BOX = 1024;  // simple_knn.cu:16 BOX_SIZE

__device__ __forceinline__ uint32_t f2ord(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// cells[0..2] = min, cells[3..5] = max, both seeded with 0 like the reference's `init = {0,0,0}`
// (simple_knn.cu:191-200).
__global__ void knn_init_cells(uint32_t* cells) {
  if (threadIdx.x < 6) cells[threadIdx.x] = f2ord(0.0f);
}

__global__ void __launch_bounds__(256)
knn_minmax_kernel(int P, const float* __restrict__ pts, uint32_t* __restrict__ cells) {
  float mn[3] = {0.f, 0.f, 0.f}, mx[3] = {0.f, 0.f, 0.f};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
#pragma unroll
    for (int c = 0; c < 3; c++) {
      float v = pts[3 * i + c];
      mn[c] = min(mn[c], v);
      mx[c] = max(mx[c], v);
    }
  }
#pragma unroll
  for (int c = 0; c < 3; c++) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn[c] = min(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], o));
      mx[c] = max(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], o));
    }
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int c = 0; c < 3; c++) {
      atomicMin(&cells[c], f2ord(mn[c]));
      atomicMax(&cells[3 + c], f2ord(mx[c]));
    }
  }
}

__device__ __forceinline__ uint32_t prep_morton(uint32_t x) {  // simple_knn.cu:45-52
  x = (x | (x << 16)) & 0x030000FF;
  x = (x | (x << 8)) & 0x0300F00F;
  x = (x | (x << 4)) & 0x030C30C3;
  x = (x | (x << 2)) & 0x09249249;
  return x;
}

__global__ void __launch_bounds__(256)
knn_morton_kernel(int P, const float* __restrict__ pts, const uint32_t* __restrict__ cells,
                  uint32_t* __restrict__ codes, uint32_t* __restrict__ ids) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  const float3 minn = {ord2f(cells[0]), ord2f(cells[1]), ord2f(cells[2])};
  const float3 maxx = {ord2f(cells[3]), ord2f(cells[4]), ord2f(cells[5])};
  const float3 c = {pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]};
  uint32_t x = prep_morton(((c.x - minn.x) / (maxx.x - minn.x)) * ((1 << 10) - 1));
  uint32_t y = prep_morton(((c.y - minn.y) / (maxx.y - minn.y)) * ((1 << 10) - 1));
  uint32_t z = prep_morton(((c.z - minn.z) / (maxx.z - minn.z)) * ((1 << 10) - 1));
  codes[i] = x | (y << 1) | (z << 2);
  ids[i] = i;
}

__global__ void __launch_bounds__(256)
knn_gather_kernel(int P, const float* __restrict__ pts, const uint32_t* __restrict__ ids,
                  float4* __restrict__ sorted) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  uint32_t id = ids[i];
  sorted[i] = {pts[3 * id], pts[3 * id + 1], pts[3 * id + 2], __uint_as_float(id)};
}

struct KnnBox { float mnx, mny, mnz, mxx, mxy, mxz, pad0, pad1; };

__global__ void __launch_bounds__(KNN_BOX)
knn_box_kernel(int P, const float4* __restrict__ sorted, KnnBox* __restrict__ boxes) {
  __shared__ float red[6][KNN_BOX / 32];
  int i = blockIdx.x * KNN_BOX + threadIdx.x;
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  if (i < P) {
    float4 p = sorted[i];
    mn[0] = mx[0] = p.x; mn[1] = mx[1] = p.y; mn[2] = mx[2] = p.z;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < 3; c++) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn[c] = min(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], o));
      mx[c] = max(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], o));
    }
    if (lane == 0) { red[c][warp] = mn[c]; red[3 + c][warp] = mx[c]; }
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int c = 0; c < 3; c++) {
      float a = red[c][lane], b = red[3 + c][lane];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        a = min(a, __shfl_xor_sync(0xffffffffu, a, o));
        b = max(b, __shfl_xor_sync(0xffffffffu, b, o));
      }
      mn[c] = a; mx[c] = b;
    }
    if (lane == 0) boxes[blockIdx.x] = {mn[0], mn[1], mn[2], mx[0], mx[1], mx[2], 0.f, 0.f};
  }
}

__device__ __forceinline__ void update_k_best(const float3& ref, const float4& point, float* knn) {
  float3 d = {point.x - ref.x, point.y - ref.y, point.z - ref.z};
  // FMA placement of the reference's SASS (boxMeanDist): fma(dz,dz, fma(dx,dx, dy*dy))
  float dist = __fmaf_rn(d.z, d.z, __fmaf_rn(d.x, d.x, __fmul_rn(d.y, d.y)));
#pragma unroll
  for (int j = 0; j < 3; j++) {
    if (knn[j] > dist) {
      float t = knn[j];
      knn[j] = dist;
      dist = t;
    }
  }
}

__device__ __forceinline__ float dist_box_point(const KnnBox& box, const float3& p) {  // simple_knn.cu:119-129
  float3 diff = {0, 0, 0};
  if (p.x < box.mnx || p.x > box.mxx) diff.x = min(abs(p.x - box.mnx), abs(p.x - box.mxx));
  if (p.y < box.mny || p.y > box.mxy) diff.y = min(abs(p.y - box.mny), abs(p.y - box.mxy));
  if (p.z < box.mnz || p.z > box.mxz) diff.z = min(abs(p.z - box.mnz), abs(p.z - box.mxz));
  return diff.x * diff.x + diff.y * diff.y + diff.z * diff.z;
}

__global__ void __launch_bounds__(256)
knn_mean_dist_kernel(int P, const float4* __restrict__ sorted, const KnnBox* __restrict__ boxes,
                     int nboxes, float* __restrict__ dists) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = idx < P;
  float4 me = valid ? sorted[idx] : make_float4(0.f, 0.f, 0.f, 0.f);
  const float3 point = {me.x, me.y, me.z};
  float best[3] = {FLT_MAX, FLT_MAX, FLT_MAX};
  if (valid) {
    for (int i = max(0, idx - 3); i <= min(P - 1, idx + 3); i++) {
      if (i == idx) continue;
      update_k_best(point, sorted[i], best);
    }
  }
  const float reject = best[2];
  best[0] = FLT_MAX; best[1] = FLT_MAX; best[2] = FLT_MAX;
  for (int b = 0; b < nboxes; b++) {
    const KnnBox box = boxes[b];
    const float dist = dist_box_point(box, point);
    const bool need = valid && !(dist > reject || dist > best[2]);
    if (!__any_sync(0xffffffffu, need)) continue;
    const int lo = b * KNN_BOX, hi = min(P, (b + 1) * KNN_BOX);
    for (int i = lo; i < hi; i++) {
      const float4 q = sorted[i];  // same address in every lane: one broadcast load
      if (need && i != idx) update_k_best(point, q, best);
    }
  }
  if (valid) dists[__float_as_uint(me.w)] = (best[0] + best[1] + best[2]) / 3.0f;
}

}  // namespace dogs

using namespace dogs;

extern "C" int sknn_dist2(int32_t P, const float* points, float* mean_dist2,
                          dogs_alloc_fn scratch_alloc, void* scratch_user, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (P <= 0) return 0;
  if (!points || !mean_dist2 || !scratch_alloc) { set_error("sknn_dist2: null pointer"); return 1; }
  const size_t n = P;
  const int nboxes = (P + KNN_BOX - 1) / KNN_BOX;
  const size_t bytes = align_up(n * 4) * 4 + align_up(n * 16) + align_up(nboxes * sizeof(KnnBox)) +
                       sort_pairs_scratch_bytes(n) + 512;
  char* sc = scratch_alloc(scratch_user, bytes);
  if (!sc) { set_error("sknn_dist2: scratch allocation failed"); return 1; }
  Carver cv(sc);
  uint32_t* codes = cv.take<uint32_t>(n);
  uint32_t* ids = cv.take<uint32_t>(n);
  uint32_t* codes_b = cv.take<uint32_t>(n);
  uint32_t* ids_b = cv.take<uint32_t>(n);
  float4* sorted = cv.take<float4>(n);
  KnnBox* boxes = cv.take<KnnBox>(nboxes);
  uint32_t* cells = cv.take<uint32_t>(8);
  char* sort_scratch = cv.take<char>(sort_pairs_scratch_bytes(n));
  knn_init_cells<<<1, 32, 0, s>>>(cells);
  int grid = (P + 255) / 256;
  knn_minmax_kernel<<<grid < 1184 ? grid : 1184, 256, 0, s>>>(P, points, cells);
  knn_morton_kernel<<<grid, 256, 0, s>>>(P, points, cells, codes, ids);
  int in_b = 0;
  if (sort_pairs_u32(codes, ids, codes_b, ids_b, n, 0, 30, sort_scratch, false, s, &in_b)) return 1;
  knn_gather_kernel<<<grid, 256, 0, s>>>(P, points, in_b ? ids_b : ids, sorted);
  knn_box_kernel<<<nboxes, KNN_BOX, 0, s>>>(P, sorted, boxes);
  knn_mean_dist_kernel<<<grid, 256, 0, s>>>(P, sorted, boxes, nboxes, mean_dist2);
  return check_launch("sknn_dist2", false, s, 6);
}
