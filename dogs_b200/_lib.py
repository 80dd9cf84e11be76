"""ctypes binding of libdogs_b200.so (the C-ABI declared in include/dogs_b200.h).

There is no CPU fallback and no alternative backend: if the shared library is missing or fails
to load, importing any operator of this package raises.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DOGS_B200_LIB") or os.path.join(_HERE, "libdogs_b200.so")     # override: A/B experiments

ALLOC_FN = C.CFUNCTYPE(C.c_void_p, C.c_void_p, C.c_size_t)


class DgrAdmmPenalty(C.Structure):
    _fields_ = [("zmu_xyz", C.c_void_p), ("zmu_dc", C.c_void_p), ("zmu_sh", C.c_void_p),
                ("zmu_scale", C.c_void_p), ("zmu_rot", C.c_void_p), ("zmu_opacity", C.c_void_p),
                ("coef_xyz", C.c_float), ("coef_dc", C.c_float), ("coef_sh", C.c_float),
                ("coef_scale", C.c_float), ("coef_rot", C.c_float), ("coef_opacity", C.c_float),
                ("sums", C.c_void_p)]


class DgrParams(C.Structure):
    _fields_ = [
        ("P", C.c_int32), ("D", C.c_int32), ("M", C.c_int32), ("W", C.c_int32), ("H", C.c_int32),
        ("tanfovx", C.c_float), ("tanfovy", C.c_float), ("scale_modifier", C.c_float),
        ("depth_threshold", C.c_float),
        ("prefiltered", C.c_int32), ("antialiasing", C.c_int32), ("debug", C.c_int32),
        ("bg", C.c_void_p), ("viewmatrix", C.c_void_p), ("projmatrix", C.c_void_p),
        ("campos", C.c_void_p),
        ("ref_bucket_cull_quirk", C.c_int32), ("raw_inputs", C.c_int32),
        ("admm", C.POINTER(DgrAdmmPenalty)),
        ("gaussians_count", C.c_void_p), ("important_score", C.c_void_p),
        ("instance_capacity", C.c_int32), ("depth_key_bits", C.c_int32),
        ("deferred_check", C.c_void_p),
    ]


class DgrForwardInfo(C.Structure):
    _fields_ = [
        ("num_rendered", C.c_int32), ("num_visible", C.c_int32),
        ("geom_bytes", C.c_size_t), ("binning_bytes", C.c_size_t),
        ("image_bytes", C.c_size_t), ("scratch_bytes", C.c_size_t),
        ("depth_key_bits", C.c_int32), ("binning_rows", C.c_int32),
    ]


class DgrLayout(C.Structure):
    _fields_ = [(n, C.c_size_t) for n in (
        "geom_record", "geom_clamped", "geom_tiles", "geom_bytes",
        "bin_point_list", "bin_tile_keys", "bin_bytes",
        "img_ranges", "img_final_T", "img_n_contrib", "img_bytes")]


class DgrAdamTensor(C.Structure):
    _fields_ = [("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
                ("M", C.c_uint32), ("lr", C.c_float), ("eps", C.c_float)]


class AdmmTensors(C.Structure):
    _fields_ = [("g", C.c_void_p * 6)]


# name -> (argtypes); every entry point returns int except the two noted below.
_P = C.c_void_p
SIGNATURES = {
    "dgr_forward": [C.POINTER(DgrParams)] + [_P] * 8 + [_P, _P, _P] + [ALLOC_FN, _P] * 4
                   + [C.POINTER(DgrForwardInfo), _P],
    "dgr_backward": [C.POINTER(DgrParams), C.c_int32] + [_P] * 8 + [_P] + [_P] * 3 + [_P, _P]
                    + [_P] * 9 + [ALLOC_FN, _P, _P],
    "dgr_visible_filter": [C.POINTER(DgrParams), _P, _P, _P, _P, _P, _P],
    "dgr_mark_visible": [C.c_int32, _P, _P, _P, _P, _P],
    "dgr_apply_activations": [C.c_int32, _P, _P, _P, _P, _P, _P, _P],
    "dgr_layout": [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(DgrLayout)],
    "dgr_deferred_check": [_P, C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)],
    "dgr_sparse_adam": [_P, _P, _P, _P, _P, C.c_float, C.c_float, C.c_float, C.c_float,
                        C.c_uint32, C.c_uint32, _P],
    "dgr_sparse_adam_multi": [C.POINTER(DgrAdamTensor), C.c_int32, _P, _P, C.c_float, C.c_float, C.c_uint32, _P],
    "dgr_compact_plan": [C.c_int64, _P, _P, ALLOC_FN, _P, C.POINTER(C.c_int64), _P],
    "dgr_compact_rows": [C.c_int64, C.c_int32, _P, _P, _P, _P, _P],
    "dgr_densify_classify": [C.c_int64, _P, _P, _P, C.c_float, C.c_float, _P, _P, ALLOC_FN, _P, C.POINTER(C.c_int64), _P],
    "dgr_densify_plan": [C.c_int64, _P, _P, C.c_int64, _P, _P, _P, _P, _P, C.c_float, C.c_float, C.c_int32, C.c_float,
                         C.c_int32, _P, _P, ALLOC_FN, _P, C.POINTER(C.c_int64), _P],
    "dgr_gather_rows": [C.c_int64, C.c_int32, _P, _P, _P, _P, C.c_int32, _P],
    "dgr_densify_split_fixup": [C.c_int64, _P, _P, _P, C.c_int64, _P, _P, _P, _P, C.c_int32, _P, _P, _P],
    "dgr_reset_opacity": [C.c_int64, _P, _P, _P, C.c_float, _P],
    "dgr_densification_stats": [C.c_int64, _P, _P, _P, _P, _P, _P],
    "fssim_forward": [C.c_int32] * 4 + [C.c_float, C.c_float] + [_P] * 6 + [_P],
    "fssim_backward": [C.c_int32] * 4 + [C.c_float, C.c_float] + [_P] * 7 + [_P],
    "fssim_l1_loss_forward": [C.c_int32] * 4 + [C.c_float] * 3 + [_P] * 5 + [_P, _P, _P],
    "fssim_l1_loss_backward": [C.c_int32] * 4 + [C.c_float] * 2 + [_P] * 5 + [_P, _P, _P],
    "sknn_dist2": [C.c_int32, _P, _P, ALLOC_FN, _P, _P],
    "admm_block_membership": [C.c_int64, _P, _P, C.c_int32, _P, _P, _P, _P],
    "admm_pack_rows": [C.POINTER(AdmmTensors), _P, C.c_int64, _P, _P],
    "admm_pack_rows_to": [C.POINTER(AdmmTensors), _P, _P, _P, C.c_int64, C.POINTER(C.c_void_p), C.c_int32, _P],
    "admm_consensus_update": [C.POINTER(AdmmTensors)] * 3 + [_P, _P, _P, _P, C.c_int32, C.c_int64,
                                                             C.POINTER(C.c_void_p), C.c_int32, C.c_float, _P, _P, _P],
    "admm_consensus_scratch_bytes": [],
    "admm_penalty": [C.POINTER(AdmmTensors)] * 3 + [_P, C.c_int64, _P, _P, _P],
    "dogs_last_error": [],
    "dogs_abi_version": [],
    "dogs_launch_count": [],
    "dogs_stage_timing": [C.c_int],
    "dogs_stage_times": [_P, C.c_int],
    "dogs_forward_stage_times": [_P, C.c_int],
}

_lib = None


def lib():
    """Loads the shared library on first use; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: build it with `python -m dogs_b200.build` "
                "(there is no CPU or PyTorch fallback for these operators)")
        _lib = C.CDLL(LIB_PATH)
        for name, args in SIGNATURES.items():
            fn = getattr(_lib, name)
            fn.argtypes = args
            fn.restype = C.c_int
        _lib.dogs_last_error.restype = C.c_char_p
        _lib.dogs_launch_count.restype = C.c_ulonglong
        _lib.admm_consensus_scratch_bytes.restype = C.c_size_t
        _lib.dogs_stage_timing.restype = None
    return _lib


def check(rc):
    if rc != 0:
        raise RuntimeError(lib().dogs_last_error().decode("utf-8", "replace"))


def ptr(t):
    """Device pointer of a tensor; None / empty tensors map to NULL (the reference's convention:
    missing optionals are empty tensors whose data_ptr is null, __init__.py:248-261)."""
    if t is None or t.numel() == 0:
        return None
    return t.data_ptr()


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def stream_ptr(device=None):
    """Raw handle of torch's current stream on `device`.  (torch.cuda.current_stream builds a Stream object and
    resolves the device through three Python layers: ~10 us, five times per training step — a tenth of the host
    time of a step on small scenes, which are host-bound.)"""
    if _raw_stream is not None and isinstance(device, torch.device) and device.index is not None:
        return _raw_stream(device.index)
    return torch.cuda.current_stream(device).cuda_stream


class _NoGuard:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


_NO_GUARD = _NoGuard()


def on_device(device):
    """`with torch.cuda.device(device)` only when `device` is not the current one already (the usual case in a
    one-process-per-GPU job: the guard's two device switches are skipped)."""
    if isinstance(device, torch.device) and device.index is not None and torch.cuda.current_device() == device.index:
        return _NO_GUARD
    return torch.cuda.device(device)


class Allocator:
    """Allocation callback backed by torch's caching allocator (mirrors resizeFunctional,
    rasterize_points.cu:30-36).  Keeps every buffer it hands out alive in `self.buffers`.

    The ctypes callback closes over the buffer list only — not over `self` — so there is no reference
    cycle: when the Allocator goes out of scope its buffers return to the caching allocator at once
    (a cycle would keep hundreds of MB per step alive until Python's cyclic GC runs, forcing fresh
    cudaMallocs and multi-millisecond stalls)."""

    def __init__(self, device):
        buffers = []
        self.device = device
        self.buffers = buffers

        def _alloc(_user, nbytes, _buffers=buffers, _device=device):
            t = torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device=_device)
            _buffers.append(t)
            return t.data_ptr()

        self.cb = ALLOC_FN(_alloc)

    @property
    def last(self):
        return self.buffers[-1] if self.buffers else None


# Allocators are reused between calls: building the ctypes callback (a closure + an ffi thunk) is a few
# microseconds each, five per training step.  An allocator goes back to the pool with an EMPTY buffer list, so the
# pool never keeps device memory alive.
_alloc_pool = {}


def take_allocators(device, n):
    pool = _alloc_pool.setdefault(device, [])
    out = []
    for _ in range(n):
        out.append(pool.pop() if pool else Allocator(device))
    return out


def give_back(allocators):
    for a in allocators:
        a.buffers.clear()
        pool = _alloc_pool.setdefault(a.device, [])
        if len(pool) < 16:
            pool.append(a)


def require_cuda(t, name):
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (dogs_b200 has no CPU path)")
    if t.dtype != torch.float32:
        raise RuntimeError(f"{name} must be float32")
    return t.contiguous()
