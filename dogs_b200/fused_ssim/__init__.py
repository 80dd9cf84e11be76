"""`fused_ssim` drop-in (API of submodules/fused-ssim/fused_ssim/__init__.py:8-41) on the sm_100a
kernels in csrc/ssim.cu.

    fused_ssim(img1, img2, padding="same", train=True) -> scalar mean SSIM
        img1, img2: [B, CH, H, W] float32 CUDA; only img1 receives a gradient.
        padding="valid" drops the 5-pixel border before averaging; train=False skips the three
        derivative maps (inference).

Low-level entry points with the reference's pybind names are exported too:
fusedssim(C1, C2, img1, img2, train) and fusedssim_backward(C1, C2, img1, img2, dL_dmap, ...).
"""
import torch

from .. import _lib as L

allowed_padding = ["same", "valid"]
_BORDER = 5                      # half width of the 11-tap window
_C1, _C2 = 0.01 ** 2, 0.03 ** 2  # SSIM stabilisers for images in [0, 1]


def _check_pair(img1, img2):
    img1 = L.require_cuda(img1, "img1")
    img2 = L.require_cuda(img2, "img2")
    if img1.dim() != 4 or img1.shape != img2.shape:
        raise RuntimeError("fusedssim expects two [B,CH,H,W] tensors of the same shape")
    return img1, img2


def fusedssim(C1, C2, img1, img2, train):
    """(ssim_map, dm_dmu1, dm_dsigma1_sq, dm_dsigma12); the last three are empty when not training
    (ssim.cu:368-404)."""
    img1, img2 = _check_pair(img1, img2)
    B, CH, H, W = img1.shape
    maps = [torch.empty_like(img1) for _ in range(4 if train else 1)]
    maps += [torch.empty(0)] * (4 - len(maps))
    if img1.numel():
        with torch.cuda.device(img1.device):
            L.check(L.lib().fssim_forward(B, CH, H, W, float(C1), float(C2), img1.data_ptr(), img2.data_ptr(),
                                          maps[0].data_ptr(), L.ptr(maps[1]), L.ptr(maps[2]), L.ptr(maps[3]),
                                          L.stream_ptr(img1.device)))
    return tuple(maps)


def fusedssim_backward(C1, C2, img1, img2, dL_dmap, dm_dmu1, dm_dsigma1_sq, dm_dsigma12):
    """dL/dimg1 from dL/dssim_map and the derivative maps of the forward (ssim.cu:406-444)."""
    img1, img2 = _check_pair(img1, img2)
    dL_dmap = L.require_cuda(dL_dmap, "dL_dmap")
    B, CH, H, W = img1.shape
    grad = torch.empty_like(img1)
    if img1.numel():
        with torch.cuda.device(img1.device):
            L.check(L.lib().fssim_backward(B, CH, H, W, float(C1), float(C2), img1.data_ptr(), img2.data_ptr(),
                                           dL_dmap.data_ptr(), dm_dmu1.data_ptr(), dm_dsigma1_sq.data_ptr(),
                                           dm_dsigma12.data_ptr(), grad.data_ptr(), L.stream_ptr(img1.device)))
    return grad


def _crop(t):
    return t[..., _BORDER:-_BORDER, _BORDER:-_BORDER]


class FusedSSIMMap(torch.autograd.Function):
    """Per-pixel SSIM map; differentiable w.r.t. img1 only (as in the reference)."""

    @staticmethod
    def forward(ctx, C1, C2, img1, img2, padding="same", train=True):
        ssim_map, *derivs = fusedssim(C1, C2, img1, img2, train)
        ctx.consts = (C1, C2, padding)
        ctx.save_for_backward(img1.detach(), img2, *derivs)
        return _crop(ssim_map) if padding == "valid" else ssim_map

    @staticmethod
    def backward(ctx, grad_map):
        C1, C2, padding = ctx.consts
        img1, img2, *derivs = ctx.saved_tensors
        if padding == "valid":            # gradient of the crop: zero border
            full = torch.zeros_like(img1)
            _crop(full).copy_(grad_map)
            grad_map = full
        grad_img1 = fusedssim_backward(C1, C2, img1, img2, grad_map, *derivs)
        return None, None, grad_img1, None, None, None


def fused_ssim(img1, img2, padding="same", train=True):
    if padding not in allowed_padding:
        raise AssertionError(f"padding must be one of {allowed_padding}")
    return FusedSSIMMap.apply(_C1, _C2, img1, img2, padding, train).mean()
