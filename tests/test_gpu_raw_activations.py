"""activations fused into preprocess (GaussianRasterizationSettings.raw_activations).

The reference activates the raw parameters in torch before every render
(conerf/model/gaussian_fields/gaussian_splat_model.py:148-153,179,202,259: sigmoid / exp /
F.normalize) and autograd chains back through them.  The fused path must give the same image and
the same gradients w.r.t. the RAW tensors as `torch activations -> rasterizer`."""
import pytest
import torch

import refutil as U
from dogs_b200 import scenes
import dogs_b200.diff_gaussian_rasterization as dgr

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GRAD_TOL = 1e-4      # BASELINE.json north_star: gradients within 1e-4 (relative to the tensor max)


def _raw_scene(cfg, seed):
    """Scene whose opacities / scales / rotations are raw parameters (inverse activations of the
    spec's values; the quaternions are deliberately left un-normalised)."""
    sc = scenes.make_scene(cfg, seed, DEV)
    g = torch.Generator(device="cpu").manual_seed(100 + seed)
    P = sc["means3D"].shape[0]
    raw = dict(sc)
    o = sc["opacities"].clamp(1e-4, 1 - 1e-4)
    raw["opacities"] = torch.log(o / (1 - o))
    raw["scales"] = torch.log(sc["scales"])
    raw["rotations"] = (sc["rotations"] * (0.25 + 3.0 * torch.rand(P, 1, generator=g).to(DEV))).contiguous()
    return raw


def _render(raw, fused, antialiasing, leaves, admm=None):
    rs = scenes.raster_settings(dgr, raw, antialiasing=antialiasing)
    if fused:
        rs = rs._replace(raw_activations=True, admm=admm)
        op, scl, rot = leaves["opacities"], leaves["scales"], leaves["rotations"]
    else:
        op = torch.sigmoid(leaves["opacities"])
        scl = torch.exp(leaves["scales"])
        rot = torch.nn.functional.normalize(leaves["rotations"])
    pts = torch.zeros_like(leaves["means3D"], requires_grad=True) + 0
    pts.retain_grad()
    img, radii, depth = dgr.GaussianRasterizer(rs)(
        means3D=leaves["means3D"], means2D=pts, dc=leaves["dc"], shs=leaves["sh"], colors_precomp=None,
        opacities=op, scales=scl, rotations=rot, cov3D_precomp=None)
    return img, radii, depth, pts, (op, scl, rot)


@pytest.mark.parametrize("cfg,antialiasing", [("tiny", False), ("tiny", True), ("C1", False)])
def test_fused_activations_match_torch_activations(cfg, antialiasing):
    raw = _raw_scene(cfg, 0)
    keys = ("means3D", "dc", "sh", "opacities", "scales", "rotations")
    g = torch.Generator(device="cpu").manual_seed(7)
    H, W = raw["H"], raw["W"]
    dLc, dLd = torch.randn(3, H, W, generator=g).to(DEV), torch.randn(1, H, W, generator=g).to(DEV)
    out = {}
    for fused in (False, True):
        leaves = {k: raw[k].clone().requires_grad_(True) for k in keys}
        img, radii, depth, pts, acts = _render(raw, fused, antialiasing, leaves)
        torch.autograd.backward([img, depth], [dLc, dLd])
        out[fused] = dict(img=img.detach(), radii=radii, depth=depth.detach(), pts=pts.grad,
                          grads={k: leaves[k].grad for k in keys}, acts=[a.detach() for a in acts])
    a, b = out[False], out[True]
    # the in-kernel activations reproduce torch's kernels bit for bit, so the forward is identical
    assert torch.equal(a["radii"], b["radii"])
    assert torch.equal(a["img"], b["img"]) and torch.equal(a["depth"], b["depth"])
    assert U.rel_to_max(b["pts"], a["pts"]) <= GRAD_TOL
    for k in keys:
        assert U.rel_to_max(b["grads"][k], a["grads"][k]) <= GRAD_TOL, k


def test_fused_activations_visible_filter_and_count():
    raw = _raw_scene("tiny", 1)
    rs = scenes.raster_settings(dgr, raw)
    act = dict(raw, opacities=torch.sigmoid(raw["opacities"]), scales=torch.exp(raw["scales"]),
               rotations=torch.nn.functional.normalize(raw["rotations"]))
    r0 = dgr.GaussianRasterizer(rs).visible_filter(act["means3D"], act["scales"], act["rotations"], None)
    r1 = dgr.GaussianRasterizer(rs._replace(raw_activations=True)).visible_filter(
        raw["means3D"], raw["scales"], raw["rotations"], None)
    assert torch.equal(r0, r1)
    c0 = dgr.rasterize_gaussians_count(act["means3D"], act["dc"], act["sh"], torch.Tensor([]), act["opacities"],
                                       act["scales"], act["rotations"], torch.Tensor([]), rs)
    c1 = dgr.rasterize_gaussians_count(raw["means3D"], raw["dc"], raw["sh"], torch.Tensor([]), raw["opacities"],
                                       raw["scales"], raw["rotations"], torch.Tensor([]),
                                       rs._replace(raw_activations=True))
    assert torch.equal(c0[0], c1[0]) and torch.equal(c0[1], c1[1])


def test_zero_quaternion_is_finite():
    """|q| below F.normalize's eps: the clamp makes the activation q / 1e-12 (= 0 for q = 0)."""
    raw = _raw_scene("tiny", 2)
    raw["rotations"][:5] = 0.0
    keys = ("means3D", "dc", "sh", "opacities", "scales", "rotations")
    leaves = {k: raw[k].clone().requires_grad_(True) for k in keys}
    img, radii, depth, pts, _ = _render(raw, True, False, leaves)
    img.sum().backward()
    assert bool(torch.isfinite(img).all())
    for k in keys:
        assert bool(torch.isfinite(leaves[k].grad).all()), k


def test_apply_activations_bit_exact_vs_torch():
    """dgr_apply_activations (the same device functions preprocess uses) against the torch kernels the
    reference calls: torch.sigmoid, torch.exp, F.normalize."""
    from dogs_b200 import _lib as L
    P = 300_001
    g = torch.Generator(device="cpu").manual_seed(3)
    op = (torch.randn(P, 1, generator=g) * 3).to(DEV)
    sc = (torch.randn(P, 3, generator=g) * 1.5 - 4).to(DEV)
    rot = (torch.randn(P, 4, generator=g) * torch.rand(P, 1, generator=g) * 3).to(DEV)
    rot[:3] = 0.0
    o, s, r = torch.empty_like(op), torch.empty_like(sc), torch.empty_like(rot)
    L.check(L.lib().dgr_apply_activations(P, op.data_ptr(), sc.data_ptr(), rot.data_ptr(), o.data_ptr(),
                                          s.data_ptr(), r.data_ptr(), L.stream_ptr(torch.device(DEV))))
    assert torch.equal(U.bits(o), U.bits(torch.sigmoid(op)))
    assert torch.equal(U.bits(s), U.bits(torch.exp(sc)))
    assert torch.equal(U.bits(r), U.bits(torch.nn.functional.normalize(rot)))


def test_raw_activations_with_all_group_admm_fusion():
    """The full fused iteration: raw parameters in, activations + all six ADMM penalty groups inside the
    rasterizer backward, against torch activations + rasterizer + stand-alone admm_penalty."""
    from dogs_b200 import admm as A
    raw = _raw_scene("tiny", 3)
    P = raw["means3D"].shape[0]
    g = torch.Generator(device="cpu").manual_seed(9)
    w = torch.randn(3, raw["H"], raw["W"], generator=g).to(DEV)
    order = ("means3D", "dc", "sh", "scales", "rotations", "opacities")

    def run(fused):
        lv = {"means3D": raw["means3D"].clone(), "dc": raw["dc"].reshape(P, 3).clone(),
              "sh": raw["sh"].reshape(P, 45).clone(), "scales": raw["scales"].clone(),
              "rotations": raw["rotations"].clone(), "opacities": raw["opacities"].clone()}
        for v in lv.values():
            v.requires_grad_(True)
        from test_gpu_aux_kernels import _virtual_plan
        blk = A.AdmmBlock([lv[k] for k in order], _virtual_plan(P, A), alphas=(3e3, 2e3, 5e3, 1e3, 7e3, 9e2))
        gg = torch.Generator(device="cpu").manual_seed(5)
        for t in blk.zmu:
            t.add_(0.05 * torch.randn(t.shape, generator=gg).to(DEV))
        blk.fuse_into_rasterizer(fused, all_groups=fused)
        leaves = dict(lv, dc=lv["dc"].view(P, 1, 3), sh=lv["sh"].view(P, 15, 3))
        img = _render(raw, fused, False, leaves, admm=blk)[0]
        (img * w).sum().backward()
        pens = blk.add_penalty_grads_().clone()
        torch.cuda.synchronize()
        return {k: v.grad.clone() for k, v in lv.items()}, pens

    g0, p0 = run(False)
    g1, p1 = run(True)
    assert torch.allclose(p0, p1, rtol=1e-6, atol=1e-15)
    for k in order:
        assert U.rel_to_max(g1[k], g0[k]) <= GRAD_TOL, k
