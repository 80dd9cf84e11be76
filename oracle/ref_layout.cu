// Test infrastructure (built into oracle/_ref/ref_layout.so by build_ref.sh, linked against the
// reference's own rasterizer_impl.o): reports the byte offsets of the fields inside the
// reference's opaque state buffers by calling its own fromChunk() carvers
// (cuda_rasterizer/rasterizer_impl.cu:249-310) on a null base pointer.  Lets the parity tests
// read n_contrib / ranges / keys and patch max_contrib without editing reference sources.
#include <cstdint>
#include <cstddef>
#include "cuda_rasterizer/rasterizer_impl.h"

using namespace CudaRasterizer;

extern "C" {
// out[0..7]: depths, clamped, internal_radii, means2D, cov3D, conic_opacity, rgb, tiles_touched; out[8] point_offsets; out[9] total
void ref_geom_offsets(size_t P, size_t* out) {
  char* p = nullptr;
  GeometryState g = GeometryState::fromChunk(p, P);
  out[0] = (size_t)g.depths; out[1] = (size_t)g.clamped; out[2] = (size_t)g.internal_radii;
  out[3] = (size_t)g.means2D; out[4] = (size_t)g.cov3D; out[5] = (size_t)g.conic_opacity;
  out[6] = (size_t)g.rgb; out[7] = (size_t)g.tiles_touched; out[8] = (size_t)g.point_offsets;
  out[9] = (size_t)p;
}
// out: point_list, point_list_unsorted, keys, keys_unsorted, total
void ref_binning_offsets(size_t R, size_t* out) {
  char* p = nullptr;
  BinningState b = BinningState::fromChunk(p, R);
  out[0] = (size_t)b.point_list; out[1] = (size_t)b.point_list_unsorted;
  out[2] = (size_t)b.point_list_keys; out[3] = (size_t)b.point_list_keys_unsorted; out[4] = (size_t)p;
}
// out: accum_alpha, n_contrib, ranges, max_contrib, pixel_colors, pixel_invDepths, bucket_count, bucket_offsets, total
void ref_image_offsets(size_t N, size_t* out) {
  char* p = nullptr;
  ImageState s = ImageState::fromChunk(p, N);
  out[0] = (size_t)s.accum_alpha; out[1] = (size_t)s.n_contrib; out[2] = (size_t)s.ranges;
  out[3] = (size_t)s.max_contrib; out[4] = (size_t)s.pixel_colors; out[5] = (size_t)s.pixel_invDepths;
  out[6] = (size_t)s.bucket_count; out[7] = (size_t)s.bucket_offsets; out[8] = (size_t)p;
}
// out: bucket_to_tile, T, ar, ard, total
void ref_sample_offsets(size_t B, size_t* out) {
  char* p = nullptr;
  SampleState s = SampleState::fromChunk(p, B);
  out[0] = (size_t)s.bucket_to_tile; out[1] = (size_t)s.T; out[2] = (size_t)s.ar; out[3] = (size_t)s.ard;
  out[4] = (size_t)p;
}
}
