#!/usr/bin/env bash
# Test infrastructure, NOT product code.
#
# Compiles the UNMODIFIED reference CUDA extensions (diff-gaussian-rasterization,
# fused-ssim, simple-knn) for sm_100a straight from the read-only sources under
# $REF (default /root/reference) into oracle/_ref/ (git-ignored, travels to the
# GPU box with gpurun).  Only build OUTPUTS land in oracle/_ref: shared objects
# and the reference's two tiny Python front-ends, copied at build time so the
# reference is driven through its own public API.  No reference source is
# committed to this repository.
#
# Numeric contract: nvcc 12.9 default flags (no -use_fast_math, -fmad default)
# plus the two include fixes recorded in SURVEY.md 8(c):
#   -include cstdint  (cuda_rasterizer/rasterizer_impl.h uses std::uintptr_t)
#   -include cfloat   (simple_knn.cu uses FLT_MAX)
set -euo pipefail
REF=${REF:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
OBJ="$OUT/obj"
PY=${PYTHON:-python}
[ -d "$REF/submodules" ] || { echo "reference not present at $REF: nothing to build"; exit 0; }
mkdir -p "$OBJ" "$OUT/diff_gaussian_rasterization" "$OUT/fused_ssim" "$OUT/simple_knn" "$OUT/old_diff_gaussian_rasterization"

read -r -a TINC <<<"$($PY - <<'PY'
from torch.utils.cpp_extension import include_paths
import sysconfig
print(" ".join("-I"+p for p in include_paths("cuda")+[sysconfig.get_paths()["include"]]))
PY
)"
TLIB="$($PY -c 'import torch,os;print(os.path.join(os.path.dirname(torch.__file__),"lib"))')"
ABI="$($PY -c 'import torch;print(int(torch._C._GLIBCXX_USE_CXX11_ABI))')"
COMMON=(-D__CUDA_NO_HALF_OPERATORS__ -D__CUDA_NO_HALF_CONVERSIONS__ -D__CUDA_NO_BFLOAT16_CONVERSIONS__
        -D__CUDA_NO_HALF2_OPERATORS__ --expt-relaxed-constexpr -std=c++17
        -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -w
        -DTORCH_API_INCLUDE_EXTENSION_H -D_GLIBCXX_USE_CXX11_ABI=$ABI "${TINC[@]}")
DGR="$REF/submodules/diff-gaussian-rasterization"
SSIM="$REF/submodules/fused-ssim"
KNN="$REF/submodules/simple-knn"
# the older rasterizer (not installed by the reference's scripts/env/install.sh, but its count mode is what
# conerf/render/gaussian_render.py:count_render and prune.py were written against): pins the LightGaussian
# importance render
OLD="$REF/submodules/old_diff-gaussian-rasterization"

cc() { # cc <obj> <src> <extname> [extra...]
  local o="$1" s="$2" n="$3"; shift 3
  if [ "$o" -nt "$s" ]; then return 0; fi
  nvcc "${COMMON[@]}" -DTORCH_EXTENSION_NAME="$n" "$@" -c "$s" -o "$o"
}
pids=()
cc "$OBJ/dgr_impl.o"    "$DGR/cuda_rasterizer/rasterizer_impl.cu" _C -I"$DGR/third_party/glm" -include cstdint & pids+=($!)
cc "$OBJ/dgr_fwd.o"     "$DGR/cuda_rasterizer/forward.cu"         _C -I"$DGR/third_party/glm" -include cstdint & pids+=($!)
cc "$OBJ/dgr_bwd.o"     "$DGR/cuda_rasterizer/backward.cu"        _C -I"$DGR/third_party/glm" -include cstdint & pids+=($!)
cc "$OBJ/dgr_adam.o"    "$DGR/cuda_rasterizer/adam.cu"            _C -I"$DGR/third_party/glm" -include cstdint & pids+=($!)
cc "$OBJ/dgr_points.o"  "$DGR/rasterize_points.cu"                _C -I"$DGR/third_party/glm" -include cstdint & pids+=($!)
cc "$OBJ/dgr_conv.o"    "$DGR/conv.cu"                            _C -I"$DGR/third_party/glm" -include cstdint & pids+=($!)
cc "$OBJ/dgr_ext.o"     "$DGR/ext.cpp"                            _C -I"$DGR/third_party/glm" -include cstdint -x cu & pids+=($!)
cc "$OBJ/ssim.o"        "$SSIM/ssim.cu"       fused_ssim_cuda & pids+=($!)
for p in "${pids[@]}"; do wait "$p"; done
pids=()
cc "$OBJ/old_impl.o"    "$OLD/cuda_rasterizer/rasterizer_impl.cu" _C -I"$OLD/third_party/glm" -include cstdint & pids+=($!)
cc "$OBJ/old_fwd.o"     "$OLD/cuda_rasterizer/forward.cu"         _C -I"$OLD/third_party/glm" -include cstdint & pids+=($!)
cc "$OBJ/old_bwd.o"     "$OLD/cuda_rasterizer/backward.cu"        _C -I"$OLD/third_party/glm" -include cstdint & pids+=($!)
cc "$OBJ/old_points.o"  "$OLD/rasterize_points.cu"                _C -I"$OLD/third_party/glm" -include cstdint & pids+=($!)
cc "$OBJ/old_ext.o"     "$OLD/ext.cpp"                            _C -I"$OLD/third_party/glm" -include cstdint -x cu & pids+=($!)
# Test-only variant of the old forward kernel with its two racy per-Gaussian counters made atomic
# (`gaussian_count[id]++; important_score[id] += opacity`, old forward.cu:485-486, are plain read-modify-writes
# executed by every pixel thread of a tile at once, so the unmodified build returns scheduling-dependent counts).
# The patched copy lives only in the git-ignored build directory.
sed -e 's/gaussian_count\[collected_id\[j\]\]++;/atomicAdd(\&gaussian_count[collected_id[j]], 1);/' \
    -e 's/important_score\[collected_id\[j\]\] += con_o.w;/atomicAdd(\&important_score[collected_id[j]], con_o.w);/' \
    "$OLD/cuda_rasterizer/forward.cu" > "$OBJ/old_fwd_atomic.cu"
grep -q 'atomicAdd(&gaussian_count' "$OBJ/old_fwd_atomic.cu" && grep -q 'atomicAdd(&important_score' "$OBJ/old_fwd_atomic.cu" \
    || { echo "old forward.cu: counter statements not found, atomic variant not built"; exit 1; }
cc "$OBJ/oldatomic_fwd.o" "$OBJ/old_fwd_atomic.cu"                _C -I"$OLD/cuda_rasterizer" -I"$OLD/third_party/glm" -include cstdint & pids+=($!)
cc "$OBJ/ssim_ext.o"    "$SSIM/ext.cpp"       fused_ssim_cuda -x cu & pids+=($!)
cc "$OBJ/knn.o"         "$KNN/simple_knn.cu"  _C -include cfloat & pids+=($!)
cc "$OBJ/knn_spatial.o" "$KNN/spatial.cu"     _C -include cfloat & pids+=($!)
cc "$OBJ/knn_ext.o"     "$KNN/ext.cpp"        _C -x cu & pids+=($!)
for p in "${pids[@]}"; do wait "$p"; done

LINK=(-shared -L"$TLIB" -lc10 -lc10_cuda -ltorch_cpu -ltorch_cuda -ltorch -ltorch_python
      -L/usr/local/cuda/lib64 -lcudart -Xlinker -rpath -Xlinker "$TLIB")
nvcc "${LINK[@]}" "$OBJ"/dgr_*.o -o "$OUT/diff_gaussian_rasterization/_C.so"
nvcc "${LINK[@]}" "$OBJ"/old_*.o -o "$OUT/old_diff_gaussian_rasterization/_C.so"
mkdir -p "$OUT/old_diff_gaussian_rasterization_atomic"
nvcc "${LINK[@]}" "$OBJ/old_impl.o" "$OBJ/oldatomic_fwd.o" "$OBJ/old_bwd.o" "$OBJ/old_points.o" "$OBJ/old_ext.o" \
     -o "$OUT/old_diff_gaussian_rasterization_atomic/_C.so"
nvcc "${LINK[@]}" "$OBJ"/ssim.o "$OBJ"/ssim_ext.o -o "$OUT/fused_ssim_cuda.so"
nvcc "${LINK[@]}" "$OBJ"/knn.o "$OBJ"/knn_spatial.o "$OBJ"/knn_ext.o -o "$OUT/simple_knn/_C.so"
# Field offsets of the reference's opaque state buffers, from its own fromChunk() carvers.
nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -Xcompiler -fPIC -w -I"$DGR" -I"$DGR/third_party/glm" \
     -include cstdint -shared "$HERE/ref_layout.cu" "$OBJ/dgr_impl.o" "$OBJ/dgr_fwd.o" "$OBJ/dgr_bwd.o" \
     -o "$OUT/ref_layout.so" -lcudart
# The reference's Python front-ends (build outputs, git-ignored): its own public API.
install -m 644 "$DGR/diff_gaussian_rasterization/__init__.py" "$OUT/diff_gaussian_rasterization/__init__.py"
install -m 644 "$SSIM/fused_ssim/__init__.py" "$OUT/fused_ssim/__init__.py"
install -m 644 "$OLD/old_diff_gaussian_rasterization/__init__.py" "$OUT/old_diff_gaussian_rasterization/__init__.py"
install -m 644 "$OLD/old_diff_gaussian_rasterization/__init__.py" "$OUT/old_diff_gaussian_rasterization_atomic/__init__.py"
: > "$OUT/simple_knn/__init__.py"
echo "reference extensions built into $OUT"
