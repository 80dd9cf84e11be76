"""Fused photometric loss (fssim_l1_loss_*): same value and gradient as the reference's
(1-l)*l1_loss + l*(1 - fused_ssim) evaluated with torch ops (gaussian_trainer.py:403-405)."""
import pytest
import torch

import refutil as U
from dogs_b200.fused_ssim import fused_ssim, fused_l1_ssim_loss

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _unfused(img, gt, lam, ssim_fn):
    l1 = (img - gt).abs().mean()
    return (1.0 - lam) * l1 + lam * (1.0 - ssim_fn(img[None], gt[None]))


@pytest.mark.parametrize("shape,lam", [((3, 64, 64), 0.2), ((3, 137, 251), 0.2), ((1, 33, 40), 0.5), ((3, 1080, 1920), 0.2)])
def test_fused_loss_matches_torch_composition(shape, lam):
    g = torch.Generator(device="cpu").manual_seed(11)
    img0 = torch.rand(*shape, generator=g).to(DEV)
    gt = torch.rand(*shape, generator=g).to(DEV)
    gt[..., :4, :4] = img0[..., :4, :4]              # exact ties: sign(0) = 0
    ssim_fns = [fused_ssim]
    if U.ref_available():
        ssim_fns.append(U.ref_modules()[1].fused_ssim)
    for fn in ssim_fns:
        a = img0.clone().requires_grad_(True)
        la = _unfused(a, gt, lam, fn)
        la.backward()
        b = img0.clone().requires_grad_(True)
        lb = fused_l1_ssim_loss(b, gt, lam)
        lb.backward()
        assert abs(la.item() - lb.item()) <= 1e-6 * max(1.0, abs(la.item()))
        # same per-pixel arithmetic as autograd's two branches; the weights are the float32 values
        # autograd produces, so the gradient agrees to the last bit or two
        assert U.rel_to_max(b.grad, a.grad) <= 1e-6


def test_fused_loss_upstream_gradient_and_repeat_calls():
    g = torch.Generator(device="cpu").manual_seed(12)
    img0 = torch.rand(3, 96, 80, generator=g).to(DEV)
    gt = torch.rand(3, 96, 80, generator=g).to(DEV)
    vals = []
    for _ in range(3):                               # the workspace must come back zeroed every time
        b = img0.clone().requires_grad_(True)
        lb = fused_l1_ssim_loss(b, gt)
        (2.5 * lb).backward()
        vals.append((lb.item(), b.grad.clone()))
    a = img0.clone().requires_grad_(True)
    (2.5 * _unfused(a, gt, 0.2, fused_ssim)).backward()
    for v, gr in vals:
        assert v == vals[0][0] or abs(v - vals[0][0]) <= 1e-7
        assert U.rel_to_max(gr, a.grad) <= 1e-6
    with torch.no_grad():                            # inference: no derivative maps
        assert abs(fused_l1_ssim_loss(img0, gt).item() - vals[0][0]) <= 1e-7
