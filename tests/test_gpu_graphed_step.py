"""GPU: a training step replayed as one CUDA graph (dogs_b200.graphs.GraphedStep) gives the eager results, and a
scene that outgrows the captured instance capacity is detected, repeated and re-captured."""
import pytest
import torch

import refutil as U
from dogs_b200 import scenes

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
KEYS = ("means3D", "dc", "sh", "opacities", "scales", "rotations")


def _make(P=3000, seed=3):
    import dogs_b200.diff_gaussian_rasterization as dgr
    from dogs_b200.fused_ssim import fused_l1_ssim_loss
    sc = scenes.make_scene("tiny", seed, DEV, P=P)
    torch.manual_seed(5)
    target = torch.rand(3, sc["H"], sc["W"], device=DEV)
    p = {k: sc[k].clone().requires_grad_(True) for k in KEYS}

    def step():
        for v in p.values():
            v.grad = None
        img, radii, depth, _ = scenes.render(dgr, sc, p)
        loss = fused_l1_ssim_loss(img, target, 0.2)
        loss.backward()
        return loss, img
    return sc, p, step


def test_graphed_step_equals_eager_step():
    from dogs_b200.graphs import GraphedStep
    sc, p, step = _make()
    loss_e, img_e = step()
    torch.cuda.synchronize()
    ge = {k: p[k].grad.clone() for k in KEYS}
    loss_e, img_e = loss_e.detach().clone(), img_e.detach().clone()
    gs = GraphedStep(step)
    for _ in range(3):
        loss_g, img_g = gs()
        assert gs.recaptures == 0
        assert torch.equal(img_g, img_e)
        assert abs(float(loss_g.detach()) - float(loss_e.detach())) <= 1e-6
        for k in KEYS:      # float atomics in the blend backward: order may differ between launches
            assert U.rel_to_max(p[k].grad, ge[k]) <= 1e-5, k


def test_graphed_step_recaptures_when_the_scene_outgrows_the_capacity():
    from dogs_b200.graphs import GraphedStep
    import dogs_b200.diff_gaussian_rasterization as dgr
    dgr._INSTANCE_CAPACITY.clear()       # what earlier tests of this process learnt
    dgr._DEPTH_KEY_BITS.clear()
    sc, p, step = _make(P=2000, seed=4)
    gs = GraphedStep(step)
    gs()
    assert gs.recaptures == 0
    with torch.no_grad():
        p["scales"].mul_(6.0)            # ~36x the footprint: the instance count leaves the 1.25x + 4096 head-room
    loss_g, img_g = gs()
    assert gs.recaptures == 1
    grads = {k: p[k].grad.clone() for k in KEYS}
    img_g, loss_g = img_g.detach().clone(), loss_g.detach().clone()
    loss_e, img_e = step()               # eager reference on the grown scene
    torch.cuda.synchronize()
    assert torch.equal(img_g, img_e) and abs(float(loss_g) - float(loss_e)) <= 1e-6
    for k in KEYS:
        assert U.rel_to_max(grads[k], p[k].grad) <= 1e-5, k
