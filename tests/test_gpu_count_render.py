"""GPU: the LightGaussian importance (count) render (GaussianRasterizationSettings.f_count, blend_fwd_kernel<true>)
against tests/golden/count_*.npz, produced on a B200 by the reference's OLDER rasterizer — the `count_gaussians`
code path conerf/render/gaussian_render.py:count_render and prune.py were written against
(tools/make_golden_count.py; oracle/build_ref.sh builds it into oracle/_ref).  Counts are integers: exact."""
import os

import numpy as np
import pytest
import torch

import dogs_b200.diff_gaussian_rasterization as dgr

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _scene(G):
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    sc = {k[3:]: t(G[k]) for k in G if k.startswith("in_")}
    sc.update(W=int(G["W"]), H=int(G["H"]), tanfovx=float(G["tanfovx"]), tanfovy=float(G["tanfovy"]))
    return sc


@pytest.mark.parametrize("name", ["count_tiny", "count_racefree"])
def test_count_render_matches_the_old_reference_rasterizer(name):
    G = dict(np.load(os.path.join(GOLD, name + ".npz")))
    sc = _scene(G)
    # the call style of the reference's count_render(): keywords, no `antialiasing`, f_count=True, fused SH tensor
    rs = dgr.GaussianRasterizationSettings(
        image_height=sc["H"], image_width=sc["W"], tanfovx=sc["tanfovx"], tanfovy=sc["tanfovy"], bg=torch.zeros(3, device=DEV),
        scale_modifier=1.0, depth_threshold=0.0, viewmatrix=sc["viewmatrix"], projmatrix=sc["projmatrix"], sh_degree=3,
        campos=sc["campos"], prefiltered=False, debug=False, f_count=True)
    cnt, score, color, radii = dgr.GaussianRasterizer(rs)(
        means3D=sc["means3D"], means2D=None, opacities=sc["opacities"], shs=torch.cat([sc["dc"], sc["sh"]], 1),
        scales=sc["scales"], rotations=sc["rotations"])
    assert np.array_equal(radii.cpu().numpy(), G["radii"])
    assert np.abs(color.cpu().numpy() - G["color"]).max() <= 1e-5
    assert cnt.dtype == torch.int32 and np.array_equal(cnt.cpu().numpy(), G["gaussians_count"])
    ref = G["important_score"]
    # the reference adds the opacity once per blended pixel in float32 (one rounding per pixel, in scheduling
    # order); here count * opacity (one rounding)
    assert np.allclose(score.cpu().numpy(), ref, rtol=1e-5, atol=1e-7)
    if name == "count_tiny":
        # what the unmodified (racy) build returned can only have lost updates
        assert bool((G["unmodified_count_run0"] <= cnt.cpu().numpy()).all())
