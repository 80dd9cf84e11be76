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


# ---- fused photometric loss (extension) ----------------------------------------------------------
_loss_ws = {}   # (device, stream) -> 32-byte workspace, zeroed once; the kernel hands it back zeroed


def _workspace(device):
    key = (device, L.stream_ptr(device))
    ws = _loss_ws.get(key)
    if ws is None:
        ws = _loss_ws[key] = torch.zeros(4, dtype=torch.float64, device=device)
    return ws


_weights_memo = {}


def _loss_weights(lambda_dssim, numel):
    """The per-pixel weights autograd would produce in float32: (-lambda) / N and (1 - lambda) / N (memoised: two
    CPU tensor divisions per step otherwise)."""
    key = (lambda_dssim, numel)
    w = _weights_memo.get(key)
    if w is None:
        n = torch.tensor(float(numel), dtype=torch.float32)
        lam = torch.tensor(lambda_dssim, dtype=torch.float32)
        if len(_weights_memo) > 64:
            _weights_memo.clear()
        w = _weights_memo[key] = (float(-lam / n), float((1.0 - lam) / n))
    return w


class _L1SSIMLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img1, img2, lambda_dssim):
        img1, img2 = _check_pair(img1, img2)
        if img1.numel() == 0:
            raise RuntimeError("fused_l1_ssim_loss: empty image")
        B, CH, H, W = img1.shape
        train = ctx.needs_input_grad[0]
        maps = [torch.empty_like(img1) for _ in range(3)] if train else [None] * 3
        loss = torch.empty((), dtype=torch.float32, device=img1.device)
        with L.on_device(img1.device):
            L.check(L.lib().fssim_l1_loss_forward(B, CH, H, W, _C1, _C2, float(lambda_dssim), img1.data_ptr(),
                                                  img2.data_ptr(), L.ptr(maps[0]), L.ptr(maps[1]), L.ptr(maps[2]),
                                                  _workspace(img1.device).data_ptr(), loss.data_ptr(),
                                                  L.stream_ptr(img1.device)))
        if train:
            ctx.save_for_backward(img1.detach(), img2, *maps)
        ctx.lambda_dssim = float(lambda_dssim)
        return loss

    @staticmethod
    def backward(ctx, grad_loss):
        img1, img2, *maps = ctx.saved_tensors
        B, CH, H, W = img1.shape
        w_ssim, w_l1 = _loss_weights(ctx.lambda_dssim, img1.numel())
        grad_loss = L.require_cuda(grad_loss, "grad_loss")
        grad = torch.empty_like(img1)
        with L.on_device(img1.device):
            L.check(L.lib().fssim_l1_loss_backward(B, CH, H, W, w_ssim, w_l1, img1.data_ptr(), img2.data_ptr(),
                                                   maps[0].data_ptr(), maps[1].data_ptr(), maps[2].data_ptr(),
                                                   grad_loss.data_ptr(), grad.data_ptr(),
                                                   L.stream_ptr(img1.device)))
        return grad, None, None


def fused_l1_ssim_loss(img1, img2, lambda_dssim=0.2):
    """(1 - lambda) * l1_loss(img1, img2) + lambda * (1 - fused_ssim(img1, img2)) — the photometric loss
    of conerf/trainers/gaussian_trainer.py:403-405 — in one forward and one backward launch instead of
    fused_ssim's two plus ~20 torch elementwise / reduction kernels.  img1, img2: [B,CH,H,W] or [CH,H,W]
    float32 CUDA; only img1 receives a gradient.  Returns a 0-d tensor."""
    if img1.dim() == 3:
        img1, img2 = img1[None], img2[None]
    return _L1SSIMLoss.apply(img1, img2, lambda_dssim)
