// Stream compaction of the Gaussian SoA (prune / densify bookkeeping).  The reference prunes with
// boolean-mask indexing on each of the 6 parameter tensors and their 12 Adam-state tensors
// (conerf/model/gaussian_fields/gaussian_splat_model.py:34-108,362-531): 18 x (nonzero + gather),
// each nonzero with its own host synchronisation.  Here the keep mask is scanned ONCE (one host
// read of the survivor count) and every tensor is gathered by a flat, coalesced kernel that reuses
// the same offsets; row order is preserved, so the result equals `t[mask]` bit for bit.
#include "common.cuh"

namespace dogs {

__global__ void __launch_bounds__(256)
mask_to_u32_kernel(const uint8_t* __restrict__ keep, uint32_t* __restrict__ out, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = keep[i] ? 1u : 0u;
}

__global__ void __launch_bounds__(256)
compact_rows_kernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, uint32_t row,
                    const uint8_t* __restrict__ keep, const uint32_t* __restrict__ offsets,
                    unsigned long long total) {
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    const unsigned long long i = e / row;
    if (keep[i]) dst[(unsigned long long)offsets[i] * row + (e - i * row)] = src[e];
  }
}

}  // namespace dogs

using namespace dogs;

// offsets[i] = number of kept rows before row i; *count_out (host) = number of kept rows.
// One stream synchronisation (the caller has to size the outputs).
extern "C" int dgr_compact_plan(int64_t P, const uint8_t* keep, uint32_t* offsets,
                                dogs_alloc_fn scratch_alloc, void* scratch_user, int64_t* count_out,
                                void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (!count_out) { set_error("compact_plan: null count_out"); return 1; }
  *count_out = 0;
  if (P <= 0) return 0;
  if (P > 0xffffffffll) { set_error("compact_plan: too many rows"); return 1; }
  if (!keep || !offsets || !scratch_alloc) { set_error("compact_plan: null pointer"); return 1; }
  const size_t n = (size_t)P;
  char* sc = scratch_alloc(scratch_user, align_up(n * 4) + scan_scratch_bytes(n) + 512);
  if (!sc) { set_error("compact_plan: scratch allocation failed"); return 1; }
  Carver cv(sc);
  uint32_t* flags = cv.take<uint32_t>(n);
  uint32_t* total = cv.take<uint32_t>(1);
  char* scan_scratch = cv.take<char>(scan_scratch_bytes(n));
  mask_to_u32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(keep, flags, (uint32_t)n);
  if (exclusive_scan_u32(flags, nullptr, offsets, total, n, scan_scratch, false, s)) return 1;
  uint32_t h = 0;
  cudaError_t e = cudaMemcpyAsync(&h, total, 4, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) { set_error("[CUDA ERROR] compact_plan: %s", cudaGetErrorString(e)); return 1; }
  *count_out = (int64_t)h;
  return check_launch("compact_plan", false, s);
}

// dst[offsets[i]] = src[i] for kept rows; rows are `row_words` 32-bit words wide (any 4-byte dtype).
extern "C" int dgr_compact_rows(int64_t P, int32_t row_words, const void* src, void* dst,
                                const uint8_t* keep, const uint32_t* offsets, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (P <= 0 || row_words <= 0) return 0;
  if (!src || !dst || !keep || !offsets) { set_error("compact_rows: null pointer"); return 1; }
  const unsigned long long total = (unsigned long long)P * (unsigned long long)row_words;
  unsigned long long g = (total + 255) / 256;
  if (g > 148ull * 32) g = 148ull * 32;
  compact_rows_kernel<<<(unsigned)g, 256, 0, s>>>((const uint32_t*)src, (uint32_t*)dst, (uint32_t)row_words,
                                                  keep, offsets, total);
  return check_launch("compact_rows", false, s);
}
