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


def test_fused_loss_and_activations_against_committed_goldens():
    """tests/golden/{loss_l1_ssim,activations}.npz were produced by the reference composition
    (tools/make_golden_loss.py); the CUDA path must reproduce them without the reference present."""
    import os
    import numpy as np
    from dogs_b200 import _lib as L
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    g = np.load(os.path.join(gold, "loss_l1_ssim.npz"))
    img = torch.from_numpy(g["img"]).to(DEV).requires_grad_(True)
    loss = fused_l1_ssim_loss(img, torch.from_numpy(g["gt"]).to(DEV), float(g["lam"]))
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) <= 1e-6
    assert U.rel_to_max(img.grad, torch.from_numpy(g["grad"]).to(DEV)) <= 1e-6
    a = np.load(os.path.join(gold, "activations.npz"))
    raw = [torch.from_numpy(a[k]).to(DEV) for k in ("raw_opacity", "raw_scaling", "raw_quaternion")]
    out = [torch.empty_like(t) for t in raw]
    L.check(L.lib().dgr_apply_activations(raw[0].shape[0], *[t.data_ptr() for t in raw], *[t.data_ptr() for t in out],
                                          L.stream_ptr(torch.device(DEV))))
    for t, k in zip(out, ("opacity", "scaling", "quaternion")):
        assert torch.equal(U.bits(t), U.bits(torch.from_numpy(a[k]).to(DEV))), k

