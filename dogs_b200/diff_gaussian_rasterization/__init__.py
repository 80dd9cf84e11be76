"""Drop-in for the reference's `diff_gaussian_rasterization` Python API
(submodules/diff-gaussian-rasterization/diff_gaussian_rasterization/__init__.py): same
`GaussianRasterizationSettings` fields, same `GaussianRasterizer.forward/markVisible/
visible_filter` signatures and return values, same `SparseGaussianAdam.step(visibility, N)`,
same validation errors — backed by the sm_100a kernels behind the C-ABI in include/dogs_b200.h.
"""
import ctypes as C
import os
from typing import NamedTuple

import torch
import torch.nn as nn

from .. import _lib as L


# False (default): exact gradients.  True: reproduce the reference's truncated backward (its
# per-tile max_contrib only covers pixel rows 0-1 of each tile; see include/dogs_b200.h
# DgrParams::ref_bucket_cull_quirk and DESIGN.md).  Also settable with DOGS_B200_REF_BWD_QUIRK=1.
REFERENCE_BACKWARD_QUIRK = os.environ.get("DOGS_B200_REF_BWD_QUIRK", "0") == "1"


class GaussianRasterizationSettings(NamedTuple):  # reference __init__.py:203-217
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool
    # The reference declares the next two without defaults; its own count_render()
    # (conerf/render/gaussian_render.py:198-213) omits `antialiasing` and passes `f_count`, which the
    # shipped rasterizer rejects with a TypeError.  Defaults + the extra f_count field make both call
    # styles work.
    antialiasing: bool = False
    depth_threshold: float = 0.0
    f_count: bool = False       # LightGaussian importance statistics (old rasterizer's count mode)
    # Extension: opacities / scales / rotations passed to the rasterizer are the model's
    # raw parameters (_opacity, _scaling, _quaternion); sigmoid / exp / F.normalize
    # (gaussian_splat_model.py:148-153) run inside preprocess and the returned gradients are w.r.t. the
    # raw tensors, replacing three torch activation kernels and their three backward kernels per render.
    raw_activations: bool = False
    # Extension: fused ADMM penalty (slave_gaussian_trainer.py:161-202).  A dogs_b200.admm.AdmmBlock (any object
    # with zmu: 6 tensors, fused_coef(): 6 floats, fused_sums: double[6] device tensor, fused_all: bool).  The
    # backward of THIS render adds coef_g * (x_g - zmu_g) to the gradients of means3D / dc / sh — and of scales /
    # rotations / opacities when the block was fused with all_groups=True, which requires raw_activations — so
    # the rasterizer must be fed the block's own parameter tensors (conerf's render() does,
    # gaussian_render.py:67-110).  Attach it to exactly one render per iteration: the penalty is added per call.
    admm: object = None


def cpu_deep_copy_tuple(input_tuple):
    return tuple(i.cpu().clone() if isinstance(i, torch.Tensor) else i for i in input_tuple)


# Capacity (in tile instances) handed to dgr_forward as DgrParams::instance_capacity, per device: 1.25x the largest
# instance count seen so far.  With a capacity the forward is enqueued in one go and the host reads the instance
# count only after the last launch (no GPU bubble); the first call on a device, and DOGS_B200_SYNC_FORWARD=1,
# use the exact, synchronising path.  A count above the capacity re-runs the binning stages inside the call.
_INSTANCE_CAPACITY = {}
# Likewise DgrParams::depth_key_bits: the largest number of significant depth-key bits seen so far (the depth sort
# runs one 8-bit pass per 8 of them; the forward verifies the assumption and repeats the stage if it was too low).
_DEPTH_KEY_BITS = {}
_SYNC_FORWARD = os.environ.get("DOGS_B200_SYNC_FORWARD", "0") == "1"


def _capacity_for(device):
    return 0 if _SYNC_FORWARD else _INSTANCE_CAPACITY.get(torch.device(device).index or 0, 0)


def _learn_capacity(device, info):
    k = torch.device(device).index or 0
    _INSTANCE_CAPACITY[k] = max(_INSTANCE_CAPACITY.get(k, 0), int(info.num_rendered * 1.25) + 4096)
    _DEPTH_KEY_BITS[k] = max(_DEPTH_KEY_BITS.get(k, 0), int(info.depth_key_bits))


# ---- deferred checks: forward calls recorded into a CUDA graph ----------------------------------------------------
# While the current stream is being captured the forward cannot wait for the instance count, so it runs in the
# library's deferred mode (DgrParams::deferred_check): the capacity and the depth-key bits learnt by earlier eager
# calls are taken for granted, the measured values land in pinned memory in stream order, and whoever replays the
# graph verifies them afterwards (dogs_b200.graphs.GraphedStep does).
class DeferredChecks:
    """Pinned slots for the forwards recorded during one capture + what each of them assumed."""

    def __init__(self, slots: int = 16):
        self.words = torch.zeros(slots, 4, dtype=torch.int32).pin_memory()
        self.entries = []          # (device index, slot, capacity, key bits)

    def take(self, device_index, capacity, key_bits):
        slot = len(self.entries)
        if slot >= self.words.shape[0]:
            raise RuntimeError("more rasterizer forwards in one captured step than DeferredChecks has slots")
        self.entries.append((device_index, slot, int(capacity), int(key_bits)))
        return self.words[slot].data_ptr()

    def ok(self):
        """After the replayed work has completed: did every recorded forward stay within what it assumed?
        Whatever the answer, the measured instance counts / key bits update the learnt capacity."""
        good = True
        for dev, slot, cap, bits in self.entries:
            r, kb = C.c_int32(0), C.c_int32(0)
            good &= bool(L.lib().dgr_deferred_check(self.words[slot].data_ptr(), cap, bits, C.byref(r), C.byref(kb)))
            _INSTANCE_CAPACITY[dev] = max(_INSTANCE_CAPACITY.get(dev, 0), int(r.value * 1.25) + 4096)
            _DEPTH_KEY_BITS[dev] = max(_DEPTH_KEY_BITS.get(dev, 0), int(kb.value))
        return good


_RECORDER = None


def begin_deferred_checks(slots: int = 16) -> DeferredChecks:
    """Call right before capturing (torch.cuda.graph) a step that contains rasterizer forwards."""
    global _RECORDER
    _RECORDER = DeferredChecks(slots)
    return _RECORDER


def end_deferred_checks():
    global _RECORDER
    _RECORDER = None


def _params(rs, P, M, device, backward=False):
    f = lambda t: t if (t.device == device and t.dtype is torch.float32 and t.is_contiguous()) \
        else t.to(device=device, dtype=torch.float32).contiguous()
    keep = (f(rs.bg), f(rs.viewmatrix), f(rs.projmatrix), f(rs.campos))
    prm = L.DgrParams(
        P=P, D=int(rs.sh_degree), M=M, W=int(rs.image_width), H=int(rs.image_height),
        tanfovx=float(rs.tanfovx), tanfovy=float(rs.tanfovy),
        scale_modifier=float(rs.scale_modifier), depth_threshold=float(rs.depth_threshold),
        prefiltered=int(bool(rs.prefiltered)), antialiasing=int(bool(rs.antialiasing)),
        debug=int(bool(rs.debug)),
        bg=keep[0].data_ptr(), viewmatrix=keep[1].data_ptr(), projmatrix=keep[2].data_ptr(),
        campos=keep[3].data_ptr(), ref_bucket_cull_quirk=int(bool(REFERENCE_BACKWARD_QUIRK)),
        raw_inputs=int(bool(getattr(rs, "raw_activations", False))))
    blk = getattr(rs, "admm", None)
    if backward and blk is not None and getattr(blk, "fused", True):
        if blk.fused_all and not getattr(rs, "raw_activations", False):
            raise RuntimeError("fused ADMM penalty with all_groups=True needs raw_activations=True: the penalty of "
                               "scaling / quaternion / opacity is taken on the tensors passed to the rasterizer, "
                               "which must be the raw consensus variables")
        zx, zd, zs = blk.zmu[0], blk.zmu[1], blk.zmu[2]
        if zx.numel() != 3 * P or zd.numel() != 3 * P or (M > 0 and zs.numel() != 3 * M * P):
            raise RuntimeError("fused ADMM penalty: the block's tensors do not match the rasterized Gaussians")
        coef = blk.fused_coef()
        pen = L.DgrAdmmPenalty(zmu_xyz=zx.data_ptr(), zmu_dc=zd.data_ptr(), zmu_sh=zs.data_ptr() if M > 0 else None,
                               coef_xyz=coef[0], coef_dc=coef[1], coef_sh=coef[2],
                               sums=blk.fused_sums.data_ptr())
        if blk.fused_all:    # scaling / quaternion / opacity tensors fed to the rasterizer ARE the consensus variables
            zc, zq, zo = blk.zmu[3], blk.zmu[4], blk.zmu[5]
            if zc.numel() != 3 * P or zq.numel() != 4 * P or zo.numel() != P:
                raise RuntimeError("fused ADMM penalty: the block's tensors do not match the rasterized Gaussians")
            pen.zmu_scale, pen.zmu_rot, pen.zmu_opacity = zc.data_ptr(), zq.data_ptr(), zo.data_ptr()
            pen.coef_scale, pen.coef_rot, pen.coef_opacity = coef[3], coef[4], coef[5]
        prm.admm = C.pointer(pen)
        keep = keep + (pen,)
    return prm, keep


def _c(t):
    return t if t.numel() == 0 else L.require_cuda(t, "input")


def rasterize_gaussians(means3D, means2D, dc, sh, colors_precomp, opacities, scales, rotations,
                        cov3Ds_precomp, raster_settings):
    return _RasterizeGaussians.apply(means3D, means2D, dc, sh, colors_precomp, opacities, scales,
                                     rotations, cov3Ds_precomp, raster_settings)


@torch.no_grad()
def rasterize_gaussians_count(means3D, dc, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                              raster_settings):
    """Count mode of the reference's older rasterizer (old_diff_gaussian_rasterization/__init__.py:
    f_count=True): returns (gaussians_count[P] int32, important_score[P] float32, color[3,H,W], radii[P]).
    Not differentiable (prune.py calls it under no_grad)."""
    rs = raster_settings
    means3D, dc, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp = map(
        _c, (means3D, dc, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp))
    device = means3D.device
    P = means3D.size(0)
    if colors_precomp.numel() == 0 and dc.numel() == 0 and sh.numel() != 0:
        dc, sh = sh[:, :1].contiguous(), sh[:, 1:].contiguous()
    M = sh.size(1) if sh.numel() != 0 else 0
    H, W = int(rs.image_height), int(rs.image_width)
    color = torch.zeros((3, H, W), dtype=torch.float32, device=device)
    invdepth = torch.empty((1, H, W), dtype=torch.float32, device=device)
    radii = torch.zeros((P,), dtype=torch.int32, device=device)
    count = torch.zeros((P,), dtype=torch.int32, device=device)
    score = torch.zeros((P,), dtype=torch.float32, device=device)
    if P != 0:
        prm, keep = _params(rs, P, M, device)
        prm.gaussians_count = count.data_ptr()
        prm.important_score = score.data_ptr()
        allocs = [L.Allocator(device) for _ in range(4)]
        info = L.DgrForwardInfo()
        with torch.cuda.device(device):
            L.check(L.lib().dgr_forward(
                C.byref(prm), L.ptr(means3D), L.ptr(dc), L.ptr(sh), L.ptr(colors_precomp), L.ptr(opacities),
                L.ptr(scales), L.ptr(rotations), L.ptr(cov3Ds_precomp), color.data_ptr(), invdepth.data_ptr(),
                radii.data_ptr(), allocs[0].cb, None, allocs[1].cb, None, allocs[2].cb, None, allocs[3].cb, None,
                C.byref(info), L.stream_ptr(device)))
        del keep
    return count, score, color, radii


class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, dc, sh, colors_precomp, opacities, scales, rotations,
                cov3Ds_precomp, raster_settings):
        rs = raster_settings
        if means3D.dim() != 2 or means3D.size(1) != 3:  # rasterize_points.cu:78-80
            raise RuntimeError("means3D must have dimensions (num_points, 3)")
        means3D, dc, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp = map(
            _c, (means3D, dc, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp))
        device = means3D.device
        P = means3D.size(0)
        # A fused [P,16,3] SH tensor without dc is accepted: split it.
        if colors_precomp.numel() == 0 and dc.numel() == 0 and sh.numel() != 0:
            dc, sh = sh[:, :1].contiguous(), sh[:, 1:].contiguous()
        M = sh.size(1) if sh.numel() != 0 else 0
        H, W = int(rs.image_height), int(rs.image_width)
        color = torch.empty((3, H, W), dtype=torch.float32, device=device)
        invdepth = torch.empty((1, H, W), dtype=torch.float32, device=device)
        radii = torch.empty((P,), dtype=torch.int32, device=device)
        geom, binning, image, scratch = L.take_allocators(device, 4)
        num_rendered = 0
        if P != 0:
            prm, keep = _params(rs, P, M, device)
            prm.instance_capacity = _capacity_for(device)
            prm.depth_key_bits = _DEPTH_KEY_BITS.get(device.index or 0, 0)
            capturing = torch.cuda.is_current_stream_capturing()
            if capturing:
                if _RECORDER is None:
                    raise RuntimeError("rasterizer forward inside a CUDA-graph capture: call begin_deferred_checks() first "
                                       "(dogs_b200.graphs.GraphedStep does all of this)")
                cap = _INSTANCE_CAPACITY.get(device.index or 0, 0)
                if cap <= 0:
                    raise RuntimeError("run the step eagerly at least once before capturing it (the instance capacity "
                                       "and the depth-key bits are learnt from eager calls)")
                prm.instance_capacity = cap
                prm.deferred_check = _RECORDER.take(device.index or 0, cap, prm.depth_key_bits)
            info = L.DgrForwardInfo()
            with L.on_device(device):
                args = (C.byref(prm), L.ptr(means3D), L.ptr(dc), L.ptr(sh), L.ptr(colors_precomp),
                        L.ptr(opacities), L.ptr(scales), L.ptr(rotations), L.ptr(cov3Ds_precomp),
                        color.data_ptr(), invdepth.data_ptr(), radii.data_ptr(),
                        geom.cb, None, binning.cb, None, image.cb, None, scratch.cb, None,
                        C.byref(info), L.stream_ptr(device))
                if rs.debug:  # reference __init__.py:92-104
                    cpu_args = cpu_deep_copy_tuple((rs.bg, means3D, colors_precomp, opacities, scales,
                                                    rotations, cov3Ds_precomp, dc, sh))
                    try:
                        L.check(L.lib().dgr_forward(*args))
                    except Exception as ex:
                        torch.save(cpu_args, "snapshot_fw.dump")
                        print("\nAn error occured in forward. Please forward snapshot_fw.dump for debugging.")
                        raise ex
                else:
                    L.check(L.lib().dgr_forward(*args))
            num_rendered = info.binning_rows    # the row count the binning buffer is carved with
            if not capturing:
                _learn_capacity(device, info)
            scratch.buffers.clear()      # stream-ordered: safe to hand back to the caching allocator now
            del keep
        else:
            color.zero_(); invdepth.zero_()
        ctx.raster_settings = rs
        ctx.num_rendered = num_rendered
        ctx.M = M
        g_buf, b_buf, i_buf = geom.last, binning.last, image.last
        if g_buf is None or b_buf is None or i_buf is None:
            empty = torch.empty(0, dtype=torch.uint8, device=device)
            g_buf = empty if g_buf is None else g_buf
            b_buf = empty if b_buf is None else b_buf
            i_buf = empty if i_buf is None else i_buf
        ctx.save_for_backward(colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, dc,
                              sh, opacities, g_buf, b_buf, i_buf)
        L.give_back((geom, binning, image, scratch))     # with empty buffer lists: ctx owns the state buffers now
        ctx.mark_non_differentiable(radii)
        # an output the loss does not use arrives as None in backward (no zero image is materialised, and
        # the blend backward compiles its depth terms out)
        ctx.set_materialize_grads(False)
        return color, radii, invdepth

    @staticmethod
    def backward(ctx, grad_out_color, _, grad_out_depth):
        rs = ctx.raster_settings
        (colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, dc, sh, opacities,
         geomBuffer, binningBuffer, imgBuffer) = ctx.saved_tensors
        device = means3D.device
        P, M = means3D.size(0), ctx.M
        new = lambda *shape: torch.empty(shape, dtype=torch.float32, device=device)
        use_sh = colors_precomp.numel() == 0
        use_cov = cov3Ds_precomp.numel() != 0
        grad_means3D, grad_means2D, grad_opacities = new(P, 3), new(P, 3), new(P, 1)
        grad_dc = new(P, 1, 3) if use_sh else None
        grad_sh = new(P, M, 3) if use_sh and M > 0 else None
        grad_colors_precomp = None if use_sh else new(P, 3)
        grad_scales = None if use_cov else new(P, 3)
        grad_rotations = None if use_cov else new(P, 4)
        grad_cov3Ds_precomp = new(P, 6) if use_cov else None
        if P != 0:
            prm, keep = _params(rs, P, M, device, backward=True)
            if grad_out_color is None:
                grad_out_color = torch.zeros((3, prm.H, prm.W), dtype=torch.float32, device=device)
            grad_out_color = grad_out_color.contiguous()
            if grad_out_depth is not None:
                grad_out_depth = grad_out_depth.contiguous()
            scratch, = L.take_allocators(device, 1)
            with L.on_device(device):
                args = (C.byref(prm), ctx.num_rendered, L.ptr(means3D), L.ptr(dc), L.ptr(sh),
                        L.ptr(colors_precomp), L.ptr(opacities), L.ptr(scales), L.ptr(rotations),
                        L.ptr(cov3Ds_precomp), radii.data_ptr(), geomBuffer.data_ptr(),
                        binningBuffer.data_ptr(), imgBuffer.data_ptr(), grad_out_color.data_ptr(),
                        L.ptr(grad_out_depth), grad_means3D.data_ptr(), grad_means2D.data_ptr(),
                        L.ptr(grad_dc), L.ptr(grad_sh), L.ptr(grad_colors_precomp),
                        grad_opacities.data_ptr(), L.ptr(grad_scales), L.ptr(grad_rotations),
                        L.ptr(grad_cov3Ds_precomp), scratch.cb, None, L.stream_ptr(device))
                if rs.debug:  # reference __init__.py:153-166
                    cpu_args = cpu_deep_copy_tuple((rs.bg, means3D, radii, colors_precomp, opacities,
                                                    scales, rotations, cov3Ds_precomp, grad_out_color,
                                                    dc, sh, grad_out_depth))
                    try:
                        L.check(L.lib().dgr_backward(*args))
                    except Exception as ex:
                        torch.save(cpu_args, "snapshot_bw.dump")
                        print("\nAn error occured in backward. Writing snapshot_bw.dump for debugging.\n")
                        raise ex
                else:
                    L.check(L.lib().dgr_backward(*args))
            L.give_back((scratch,))      # stream-ordered: the accumulator may return to the caching allocator now
            del keep
        # depth_threshold scaling of grad_means2D (reference __init__.py:171-185) is fused into
        # the preprocess-backward kernel.
        return (grad_means3D, grad_means2D, grad_dc, grad_sh, grad_colors_precomp, grad_opacities,
                grad_scales, grad_rotations, grad_cov3Ds_precomp, None)


class GaussianRasterizer(nn.Module):  # reference __init__.py:220-300
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions):
        with torch.no_grad():
            rs = self.raster_settings
            positions = L.require_cuda(positions, "positions")
            P = positions.size(0)
            visible = torch.empty((P,), dtype=torch.bool, device=positions.device)
            if P:
                view = rs.viewmatrix.to(positions.device, torch.float32).contiguous()
                proj = rs.projmatrix.to(positions.device, torch.float32).contiguous()
                with torch.cuda.device(positions.device):
                    L.check(L.lib().dgr_mark_visible(P, positions.data_ptr(), view.data_ptr(),
                                                     proj.data_ptr(), visible.data_ptr(),
                                                     L.stream_ptr(positions.device)))
        return visible

    def forward(self, means3D, means2D, opacities, dc=None, shs=None, colors_precomp=None,
                scales=None, rotations=None, cov3D_precomp=None):
        raster_settings = self.raster_settings
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')
        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception(
                'Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
        if dc is None:
            dc = torch.Tensor([])
        if shs is None:
            shs = torch.Tensor([])
        if colors_precomp is None:
            colors_precomp = torch.Tensor([])
        if scales is None:
            scales = torch.Tensor([])
        if rotations is None:
            rotations = torch.Tensor([])
        if cov3D_precomp is None:
            cov3D_precomp = torch.Tensor([])
        if getattr(raster_settings, "f_count", False):
            return rasterize_gaussians_count(means3D, dc, shs, colors_precomp, opacities, scales, rotations,
                                             cov3D_precomp, raster_settings)
        return rasterize_gaussians(means3D, means2D, dc, shs, colors_precomp, opacities, scales,
                                   rotations, cov3D_precomp, raster_settings)

    def visible_filter(self, means3D, scales=None, rotations=None, cov3D_precomp=None):
        rs = self.raster_settings
        with torch.no_grad():
            means3D = L.require_cuda(means3D, "means3D")
            device = means3D.device
            P = means3D.size(0)
            radii = torch.empty((P,), dtype=torch.int32, device=device)
            if P:
                prm, keep = _params(rs, P, 0, device)
                sc = None if scales is None else _c(scales)
                ro = None if rotations is None else _c(rotations)
                cv = None if cov3D_precomp is None else _c(cov3D_precomp)
                with torch.cuda.device(device):
                    L.check(L.lib().dgr_visible_filter(C.byref(prm), means3D.data_ptr(), L.ptr(sc),
                                                       L.ptr(ro), L.ptr(cv), radii.data_ptr(),
                                                       L.stream_ptr(device)))
                del keep
        return radii


class SparseGaussianAdam(torch.optim.Adam):  # reference __init__.py:303-332
    def __init__(self, params, lr, eps):
        super().__init__(params=params, lr=lr, eps=eps)

    @torch.no_grad()
    def step(self, visibility, N):
        """visibility: bool [N] (the reference's `radii > 0`, gaussian_trainer.py:474-475) or the rasterizer's int32
        radii themselves.  All param groups are updated by ONE kernel launch (dgr_sparse_adam_multi) where the
        reference launches adamUpdate once per group; the arithmetic per element is the same."""
        todo = []
        for group in self.param_groups:
            lr = group["lr"]
            eps = group["eps"]
            assert len(group["params"]) == 1, "more than one tensor in group"
            param = group["params"][0]
            if param.grad is None:
                continue
            state = self.state[param]
            if len(state) == 0:
                state['step'] = torch.tensor(0.0, dtype=torch.float32)
                state['exp_avg'] = torch.zeros_like(param, memory_format=torch.preserve_format)
                state['exp_avg_sq'] = torch.zeros_like(param, memory_format=torch.preserve_format)
            todo.append((param, param.grad, state["exp_avg"], state["exp_avg_sq"], lr, eps))
        adamUpdateMulti(todo, visibility, 0.9, 0.999, N)


def adamUpdateMulti(tensors, visible, b1, b2, N):
    """tensors: list of (param, grad, exp_avg, exp_avg_sq, lr, eps); in-place sparse Adam step on all of them."""
    if not tensors or N == 0:
        return
    dev = tensors[0][0].device
    vis = visible.contiguous()
    if vis.dtype not in (torch.bool, torch.int32) or vis.numel() != N:
        raise RuntimeError("visibility must be a bool tensor, or the int32 radii, of N elements")
    keep = []
    for lo in range(0, len(tensors), 8):
        chunk = tensors[lo:lo + 8]
        arr = (L.DgrAdamTensor * len(chunk))()
        for j, (param, grad, m, v, lr, eps) in enumerate(chunk):
            for t in (param, m, v):
                if not (t.is_cuda and t.is_contiguous() and t.dtype == torch.float32):
                    raise RuntimeError("adamUpdate needs contiguous float32 CUDA tensors")
            g = grad.contiguous()
            keep.append(g)
            if param.numel() % N:
                raise RuntimeError("parameter size is not a multiple of N")
            arr[j] = L.DgrAdamTensor(param.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), param.numel() // N,
                                     float(lr), float(eps))
        with torch.cuda.device(dev):
            L.check(L.lib().dgr_sparse_adam_multi(arr, len(chunk), vis.data_ptr() if vis.dtype == torch.bool else None,
                                                  vis.data_ptr() if vis.dtype == torch.int32 else None, float(b1),
                                                  float(b2), int(N), L.stream_ptr(dev)))


def adamUpdate(param, param_grad, exp_avg, exp_avg_sq, visible, lr, b1, b2, eps, N, M):
    """_C.adamUpdate (rasterize_points.cu:331-361): in-place sparse Adam step."""
    for t in (param, exp_avg, exp_avg_sq):
        if not (t.is_cuda and t.is_contiguous() and t.dtype == torch.float32):
            raise RuntimeError("adamUpdate needs contiguous float32 CUDA tensors")
    grad = param_grad.contiguous()
    vis = visible.contiguous()
    if vis.dtype != torch.bool:
        raise RuntimeError("visibility must be a bool tensor")
    if N * M == 0:
        return
    with torch.cuda.device(param.device):
        L.check(L.lib().dgr_sparse_adam(param.data_ptr(), grad.data_ptr(), exp_avg.data_ptr(),
                                        exp_avg_sq.data_ptr(), vis.data_ptr(), float(lr), float(b1),
                                        float(b2), float(eps), int(N), int(M),
                                        L.stream_ptr(param.device)))
