"""Drop-in for the reference's `fused_ssim` package (submodules/fused-ssim/fused_ssim/__init__.py):
`fused_ssim(img1, img2, padding="same", train=True)` -> mean SSIM, differentiable w.r.t. img1 only."""
import torch

from .. import _lib as L

allowed_padding = ["same", "valid"]


def fusedssim(C1, C2, img1, img2, train):
    """fused_ssim_cuda.fusedssim (ssim.cu:368-404) -> (ssim_map, dm_dmu1, dm_dsigma1_sq, dm_dsigma12)."""
    img1 = L.require_cuda(img1, "img1")
    img2 = L.require_cuda(img2, "img2")
    if img1.dim() != 4 or img1.shape != img2.shape:
        raise RuntimeError("fusedssim expects two [B,CH,H,W] tensors of the same shape")
    B, CH, H, W = img1.shape
    ssim_map = torch.empty_like(img1)
    if train:
        d1, d2, d3 = torch.empty_like(img1), torch.empty_like(img1), torch.empty_like(img1)
    else:
        d1 = d2 = d3 = torch.empty(0)
    if img1.numel():
        with torch.cuda.device(img1.device):
            L.check(L.lib().fssim_forward(B, CH, H, W, float(C1), float(C2), img1.data_ptr(),
                                          img2.data_ptr(), ssim_map.data_ptr(), L.ptr(d1), L.ptr(d2),
                                          L.ptr(d3), L.stream_ptr(img1.device)))
    return ssim_map, d1, d2, d3


def fusedssim_backward(C1, C2, img1, img2, dL_dmap, dm_dmu1, dm_dsigma1_sq, dm_dsigma12):
    """fused_ssim_cuda.fusedssim_backward (ssim.cu:406-444) -> dL_dimg1."""
    img1 = L.require_cuda(img1, "img1")
    img2 = L.require_cuda(img2, "img2")
    dL_dmap = L.require_cuda(dL_dmap, "dL_dmap")
    B, CH, H, W = img1.shape
    out = torch.empty_like(img1)
    if img1.numel():
        with torch.cuda.device(img1.device):
            L.check(L.lib().fssim_backward(B, CH, H, W, float(C1), float(C2), img1.data_ptr(),
                                           img2.data_ptr(), dL_dmap.data_ptr(), dm_dmu1.data_ptr(),
                                           dm_dsigma1_sq.data_ptr(), dm_dsigma12.data_ptr(),
                                           out.data_ptr(), L.stream_ptr(img1.device)))
    return out


class FusedSSIMMap(torch.autograd.Function):
    @staticmethod
    def forward(ctx, C1, C2, img1, img2, padding="same", train=True):
        ssim_map, dm_dmu1, dm_dsigma1_sq, dm_dsigma12 = fusedssim(C1, C2, img1, img2, train)
        if padding == "valid":
            ssim_map = ssim_map[:, :, 5:-5, 5:-5]
        ctx.save_for_backward(img1.detach(), img2, dm_dmu1, dm_dsigma1_sq, dm_dsigma12)
        ctx.C1 = C1
        ctx.C2 = C2
        ctx.padding = padding
        return ssim_map

    @staticmethod
    def backward(ctx, opt_grad):
        img1, img2, dm_dmu1, dm_dsigma1_sq, dm_dsigma12 = ctx.saved_tensors
        C1, C2, padding = ctx.C1, ctx.C2, ctx.padding
        dL_dmap = opt_grad
        if padding == "valid":
            dL_dmap = torch.zeros_like(img1)
            dL_dmap[:, :, 5:-5, 5:-5] = opt_grad
        grad = fusedssim_backward(C1, C2, img1, img2, dL_dmap, dm_dmu1, dm_dsigma1_sq, dm_dsigma12)
        return None, None, grad, None, None, None


def fused_ssim(img1, img2, padding="same", train=True):
    C1 = 0.01 ** 2
    C2 = 0.03 ** 2
    assert padding in allowed_padding
    map = FusedSSIMMap.apply(C1, C2, img1, img2, padding, train)
    return map.mean()
