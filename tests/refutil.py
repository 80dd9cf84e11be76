"""Helpers shared by the GPU parity tests: import the reference build (oracle/_ref, test
infrastructure only), parse its opaque state buffers, compare results.

Buffer layouts follow CudaRasterizer::{GeometryState,BinningState,ImageState}::fromChunk
(cuda_rasterizer/rasterizer_impl.cu:249-310): fields carved in order, each aligned to 128 B."""
import ctypes as C
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")


def ref_available():
    return os.path.exists(os.path.join(REF_DIR, "diff_gaussian_rasterization", "_C.so"))


_cache = {}


def ref_modules():
    """(diff_gaussian_rasterization, fused_ssim, simple_knn._C) of the UNMODIFIED reference."""
    if "mods" not in _cache:
        saved = {k: sys.modules.pop(k) for k in list(sys.modules)
                 if k.split(".")[0] in ("diff_gaussian_rasterization", "fused_ssim", "simple_knn", "fused_ssim_cuda")}
        sys.path.insert(0, REF_DIR)
        try:
            dgr = importlib.import_module("diff_gaussian_rasterization")
            ssim = importlib.import_module("fused_ssim")
            knn = importlib.import_module("simple_knn._C")
        finally:
            sys.path.remove(REF_DIR)
            for k in list(sys.modules):
                if k.split(".")[0] in ("diff_gaussian_rasterization", "fused_ssim", "simple_knn", "fused_ssim_cuda"):
                    sys.modules.pop(k)
            sys.modules.update(saved)
        _cache["mods"] = (dgr, ssim, knn)
    return _cache["mods"]


_layout = None


def _ref_layout():
    global _layout
    if _layout is None:
        _layout = C.CDLL(os.path.join(REF_DIR, "ref_layout.so"))
    return _layout


def _offsets(fn, n, k):
    out = (C.c_size_t * 16)()
    getattr(_ref_layout(), fn)(C.c_size_t(n), out)
    return list(out)[:k]


def _view(buf, off, dtype, count):
    nbytes = count * torch.empty(0, dtype=dtype).element_size()
    base = buf.data_ptr()
    assert base % 128 == 0, "reference state buffer is not 128-byte aligned"
    return buf[off:off + nbytes].view(dtype)


def parse_ref_geom(buf, P):
    o = _offsets("ref_geom_offsets", P, 10)
    names = [("depths", torch.float32, P), ("clamped", torch.uint8, 3 * P),
             ("internal_radii", torch.int32, P), ("means2D", torch.float32, 2 * P),
             ("cov3D", torch.float32, 6 * P), ("conic_opacity", torch.float32, 4 * P),
             ("rgb", torch.float32, 3 * P), ("tiles_touched", torch.int32, P),
             ("point_offsets", torch.int32, P)]
    return {n: _view(buf, o[i], dt, c) for i, (n, dt, c) in enumerate(names)}


def parse_ref_binning(buf, R):
    o = _offsets("ref_binning_offsets", R, 5)
    names = [("point_list", torch.int32, R), ("point_list_unsorted", torch.int32, R),
             ("keys", torch.int64, R), ("keys_unsorted", torch.int64, R)]
    return {n: _view(buf, o[i], dt, c) for i, (n, dt, c) in enumerate(names)}


def parse_ref_image(buf, N):
    o = _offsets("ref_image_offsets", N, 9)
    names = [("accum_alpha", torch.float32, N), ("n_contrib", torch.int32, N),
             ("ranges", torch.int32, 2 * N), ("max_contrib", torch.int32, N),
             ("pixel_colors", torch.float32, 3 * N), ("pixel_invDepths", torch.float32, N),
             ("bucket_count", torch.int32, N), ("bucket_offsets", torch.int32, N)]
    return {n: _view(buf, o[i], dt, c) for i, (n, dt, c) in enumerate(names)}


def tile_max(n_contrib, H, W, rows=16):
    """Per-tile max of n_contrib over the first `rows` pixel rows of each 16x16 tile -> int32 [T]."""
    gy, gx = (H + 15) // 16, (W + 15) // 16
    pad = torch.zeros(gy * 16, gx * 16, dtype=torch.int32, device=n_contrib.device)
    pad[:H, :W] = n_contrib.view(H, W)
    t = pad.view(gy, 16, gx, 16)[:, :rows]
    return t.amax(dim=(1, 3)).reshape(-1).contiguous()


def ref_backward_raw(dgr, scene, fwd, dL_dcolor, dL_dinvdepth, bg=None, antialiasing=False):
    """The reference's pybind backward on the buffers of ref_forward_raw().  Returns a dict of the
    gradients in the reference's shapes (rasterize_points.cu:156-252)."""
    dev = scene["means3D"].device
    if bg is None:
        bg = torch.zeros(3, device=dev)
    e = torch.Tensor([])
    out = dgr._C.rasterize_gaussians_backward(
        bg, scene["means3D"], fwd["radii"], e, scene["opacities"], scene["scales"], scene["rotations"],
        1.0, e, scene["viewmatrix"], scene["projmatrix"], scene["tanfovx"], scene["tanfovy"], dL_dcolor,
        scene["dc"], scene["sh"], dL_dinvdepth, int(scene["sh_degree"]), scene["campos"], fwd["geom"],
        fwd["num_rendered"], fwd["binning"], fwd["img"], fwd["num_buckets"], fwd["sample"], antialiasing,
        False)
    keys = ("means2D", "colors", "opacities", "means3D", "cov3D", "dc", "sh", "scales", "rotations", "depth")
    return dict(zip(keys, out))


def ours_backward(scene, dL_dcolor, dL_dinvdepth, quirk=False, antialiasing=False, depth_threshold=0.0):
    """Forward + backward of this package through its public autograd API with given output grads."""
    import dogs_b200.diff_gaussian_rasterization as mod
    from dogs_b200 import scenes
    old = mod.REFERENCE_BACKWARD_QUIRK
    mod.REFERENCE_BACKWARD_QUIRK = quirk
    try:
        p = {k: scene[k].clone().requires_grad_(True)
             for k in ("means3D", "dc", "sh", "opacities", "scales", "rotations")}
        img, radii, depth, spts = scenes.render(mod, scene, p, antialiasing=antialiasing,
                                                depth_threshold=depth_threshold)
        torch.autograd.backward([img, depth], [dL_dcolor, dL_dinvdepth])
        torch.cuda.synchronize()
    finally:
        mod.REFERENCE_BACKWARD_QUIRK = old
    g = {k: v.grad for k, v in p.items()}
    g["means2D"] = spts.grad
    return g


def ref_forward_raw(dgr, scene, bg=None, antialiasing=False, debug=False):
    """Calls the reference's pybind entry directly to get its state buffers."""
    dev = scene["means3D"].device
    if bg is None:
        bg = torch.zeros(3, device=dev)
    e = torch.Tensor([])
    out = dgr._C.rasterize_gaussians(
        bg, scene["means3D"], e, scene["opacities"], scene["scales"], scene["rotations"], 1.0, e,
        scene["viewmatrix"], scene["projmatrix"], scene["tanfovx"], scene["tanfovy"], int(scene["H"]),
        int(scene["W"]), scene["dc"], scene["sh"], int(scene["sh_degree"]), scene["campos"], False,
        antialiasing, debug)
    keys = ("num_rendered", "num_buckets", "color", "invdepth", "radii", "geom", "binning", "img", "sample")
    return dict(zip(keys, out))


def ours_forward_raw(scene, bg=None, antialiasing=False, depth_threshold=0.0, debug=False, instance_capacity=0,
                     depth_key_bits=0):
    """Calls dgr_forward through the C-ABI and returns outputs + parsed state.  instance_capacity > 0 selects
    the path that enqueues the whole pass before the host knows the instance count."""
    from dogs_b200 import _lib as L
    from dogs_b200.diff_gaussian_rasterization import _params, GaussianRasterizationSettings
    dev = scene["means3D"].device
    if bg is None:
        bg = torch.zeros(3, device=dev)
    P, H, W = scene["means3D"].shape[0], int(scene["H"]), int(scene["W"])
    rs = GaussianRasterizationSettings(H, W, scene["tanfovx"], scene["tanfovy"], bg, 1.0,
                                       scene["viewmatrix"], scene["projmatrix"], int(scene["sh_degree"]),
                                       scene["campos"], False, debug, antialiasing, depth_threshold)
    M = scene["sh"].shape[1]
    prm, keep = _params(rs, P, M, dev)
    prm.instance_capacity = int(instance_capacity)
    prm.depth_key_bits = int(depth_key_bits)
    color = torch.empty(3, H, W, device=dev)
    invdepth = torch.empty(1, H, W, device=dev)
    radii = torch.empty(P, dtype=torch.int32, device=dev)
    al = [L.Allocator(dev) for _ in range(4)]
    info = L.DgrForwardInfo()
    L.check(L.lib().dgr_forward(C.byref(prm), scene["means3D"].data_ptr(), scene["dc"].data_ptr(),
                                scene["sh"].data_ptr(), None, scene["opacities"].data_ptr(),
                                scene["scales"].data_ptr(), scene["rotations"].data_ptr(), None,
                                color.data_ptr(), invdepth.data_ptr(), radii.data_ptr(),
                                al[0].cb, None, al[1].cb, None, al[2].cb, None, al[3].cb, None,
                                C.byref(info), L.stream_ptr(dev)))
    torch.cuda.synchronize()
    R = info.num_rendered
    lay = L.DgrLayout()
    L.check(L.lib().dgr_layout(P, W, H, info.binning_rows, C.byref(lay)))
    g, b, im = al[0].last, al[1].last, al[2].last
    N = H * W
    T = ((W + 15) // 16) * ((H + 15) // 16)
    rec = g[lay.geom_record:lay.geom_record + 48 * P].view(torch.float32).view(P, 12)
    out = dict(num_rendered=R, binning_rows=info.binning_rows, depth_key_bits=info.depth_key_bits, color=color, invdepth=invdepth, radii=radii,
               conic_opacity=rec[:, 0:4], means2D=rec[:, 4:6], invdepths=rec[:, 6],
               rgb=torch.stack([rec[:, 7], rec[:, 8], rec[:, 9]], 1), depths=rec[:, 10],
               clamped=g[lay.geom_clamped:lay.geom_clamped + P],
               tiles=g[lay.geom_tiles:lay.geom_tiles + 4 * P].view(torch.int32),
               point_list=b[lay.bin_point_list:lay.bin_point_list + 4 * R].view(torch.int32),
               tile_keys=b[lay.bin_tile_keys:lay.bin_tile_keys + 4 * R].view(torch.int32),
               ranges=im[lay.img_ranges:lay.img_ranges + 8 * T].view(torch.int32).view(T, 2),
               final_T=im[lay.img_final_T:lay.img_final_T + 4 * N].view(torch.float32),
               n_contrib=im[lay.img_n_contrib:lay.img_n_contrib + 4 * N].view(torch.int32),
               _keep=(al, keep), _bg=bg, _rec=g[lay.geom_record:lay.geom_record + 48 * P])
    return out


_truth = None


def truth_available():
    return os.path.exists(os.path.join(ROOT, "oracle", "libtruth.so"))


def truth_backward(scene, o, dL_dcolor, dL_dinvdepth, depth_threshold=0.0):
    """FP64 evaluation of the reference's backward formulas (oracle/truth_bwd.cu, test infrastructure) on the
    float32 forward state `o` = ours_forward_raw(scene), which is bit-identical to the reference's.  Returns
    float64 gradients keyed like GRAD_KEYS; no antialiasing, scale/rotation inputs, SH colours."""
    global _truth
    if _truth is None:
        _truth = C.CDLL(os.path.join(ROOT, "oracle", "libtruth.so"))
    dev = scene["means3D"].device
    P, H, W = scene["means3D"].shape[0], int(scene["H"]), int(scene["W"])
    M, D = scene["sh"].shape[1], int(scene["sh_degree"])
    acc = torch.zeros(P, 10, dtype=torch.float64, device=dev)
    p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
    dLc = dL_dcolor.contiguous()
    dLd = dL_dinvdepth.contiguous() if dL_dinvdepth is not None else None
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    rc = _truth.truth_blend_bwd(C.c_int(W), C.c_int(H), p(o["ranges"]), p(o["point_list"]), p(o["_bg"]), p(o["_rec"]),
                                p(o["n_contrib"]), p(dLc), p(dLd), p(acc), st)
    assert rc == 0, f"truth_blend_bwd: CUDA error {rc}"
    new = lambda *s: torch.empty(s, dtype=torch.float64, device=dev)
    g3, g2, gdc, gsh, gop, gsc, grot = new(P, 3), new(P, 2), new(P, 3), new(P, M, 3), new(P), new(P, 3), new(P, 4)
    f32 = lambda v: float(torch.tensor(v, dtype=torch.float32))      # the forward consumed float32 values
    rc = _truth.truth_preprocess_bwd(
        C.c_int(P), C.c_int(D), C.c_int(M), C.c_int(W), C.c_int(H), p(scene["means3D"]), p(o["radii"]), p(scene["sh"]),
        p(o["clamped"]), p(scene["scales"]), p(scene["rotations"]), C.c_double(1.0), p(scene["viewmatrix"]),
        p(scene["projmatrix"]), C.c_double(f32(scene["tanfovx"])), C.c_double(f32(scene["tanfovy"])), p(scene["campos"]),
        p(acc), C.c_double(depth_threshold), p(g3), p(g2), p(gdc), p(gsh), p(gop), p(gsc), p(grot), st)
    assert rc == 0, f"truth_preprocess_bwd: CUDA error {rc}"
    torch.cuda.synchronize()
    m2 = torch.zeros(P, 3, dtype=torch.float64, device=dev)
    m2[:, :2] = g2
    return {"means3D": g3, "means2D": m2, "dc": gdc.view(P, 1, 3), "sh": gsh, "opacities": gop.view(P, 1),
            "scales": gsc, "rotations": grot}


def err_vs_truth(a, truth):
    """max |a - truth| / max |truth| with the comparison done in float64."""
    t = truth.reshape(a.shape)
    return float((a.double() - t).abs().max() / t.abs().max().clamp_min(1e-300))


def bits(t):
    return t.contiguous().view(torch.int32)


def rel_to_max(a, b):
    """max |a-b| / max |b|  (per-tensor relative error; SURVEY.md 7 'backward nondeterminism budget')."""
    denom = b.abs().max().item()
    if denom == 0:
        return (a - b).abs().max().item()
    return (a - b).abs().max().item() / denom


def count_outliers(a, b, tol):
    """number of elements with |a-b| > tol * max|b|"""
    denom = b.abs().max().item()
    return int(((a - b).abs() > tol * max(denom, 1e-30)).sum().item())

