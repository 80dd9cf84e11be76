"""Golden vectors for the LightGaussian importance (count) render, produced ON A B200 by the reference's OLDER
rasterizer (submodules/old_diff-gaussian-rasterization, `count_gaussians`; built by oracle/build_ref.sh into
oracle/_ref/old_diff_gaussian_rasterization) — the code conerf/render/gaussian_render.py:count_render and
prune.py were written against.

Its two per-Gaussian counters are plain `x[id]++` / `x[id] += opacity` executed by all pixel threads of a tile
at once (old forward.cu:485-486), so the unmodified build returns scheduling-dependent numbers whenever a
Gaussian covers more than one pixel.  Two pins are stored:
  count_tiny      the ordinary tiny scene through the test-only build whose two statements are atomicAdd
                  (oracle/build_ref.sh, ..._atomic) + colour / radii of the UNMODIFIED build;
  count_racefree  a scene of sub-pixel, nearly transparent Gaussians (each reaches alpha >= 1/255 on at most one
                  pixel), where the UNMODIFIED build is race-free and both builds must agree.

    python tools/make_golden_count.py     (where oracle/_ref exists, on a GPU)  ->  tests/golden/count_*.npz
"""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.path.join(ROOT, "oracle", "_ref")
sys.path.insert(0, REF)
from dogs_b200 import scenes  # noqa: E402

DEV = "cuda:0"
old = importlib.import_module("old_diff_gaussian_rasterization")
old_atomic = importlib.import_module("old_diff_gaussian_rasterization_atomic")
npy = lambda t: t.detach().cpu().numpy()


def run(mod, sc):
    rs = mod.GaussianRasterizationSettings(
        image_height=int(sc["H"]), image_width=int(sc["W"]), tanfovx=sc["tanfovx"], tanfovy=sc["tanfovy"],
        bg=torch.zeros(3, device=DEV), scale_modifier=1.0, viewmatrix=sc["viewmatrix"], projmatrix=sc["projmatrix"],
        sh_degree=3, campos=sc["campos"], prefiltered=False, debug=False, depth_threshold=0.0, f_count=True)
    shs = torch.cat([sc["dc"], sc["sh"]], dim=1).contiguous()
    cnt, score, color, radii = mod.GaussianRasterizer(rs)(
        means3D=sc["means3D"], means2D=torch.zeros_like(sc["means3D"]), opacities=sc["opacities"], shs=shs,
        scales=sc["scales"], rotations=sc["rotations"])
    torch.cuda.synchronize()
    return cnt, score, color, radii


def save(name, sc, cnt, score, color, radii, **extra):
    out = {f"in_{k}": npy(sc[k]) for k in ("means3D", "dc", "sh", "opacities", "scales", "rotations", "viewmatrix",
                                           "projmatrix", "campos")}
    out.update(W=sc["W"], H=sc["H"], tanfovx=sc["tanfovx"], tanfovy=sc["tanfovy"], sh_degree=3,
               gaussians_count=npy(cnt), important_score=npy(score), color=npy(color), radii=npy(radii), **extra)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), **out)


sc = scenes.make_scene("tiny", 12, DEV)
c_a, s_a, col_a, rad_a = run(old_atomic, sc)
runs = [run(old, sc) for _ in range(3)]
c_u, s_u, col_u, rad_u = runs[0]
assert torch.equal(col_a, col_u) and torch.equal(rad_a, rad_u)
racy = [int((r[0] != c_a).sum()) for r in runs]
assert all(bool((r[0] <= c_a).all()) for r in runs)        # lost updates only
print("tiny: atomic count sum", int(c_a.sum()), "| unmodified build differs on", racy, "of", c_a.numel(), "Gaussians")
save("count_tiny", sc, c_a, s_a, col_u, rad_u, unmodified_count_run0=npy(c_u), n_gaussians_racy=np.array(racy))

g = torch.Generator(device="cpu").manual_seed(3)
sc2 = scenes.make_scene("tiny", 13, DEV, P=3000)
sc2["scales"] = torch.full_like(sc2["scales"], 1e-5)
sc2["opacities"] = (0.0042 + 0.0016 * torch.rand(3000, 1, generator=g)).to(DEV)
c2a, s2a, col2a, rad2a = run(old_atomic, sc2)
for _ in range(3):
    c2, s2, col2, rad2 = run(old, sc2)
    assert torch.equal(c2, c2a) and torch.equal(s2, s2a) and torch.equal(col2, col2a)
assert int(c2.max()) == 1 and int(c2.sum()) > 100
print("racefree: count sum", int(c2.sum()), "max", int(c2.max()))
save("count_racefree", sc2, c2, s2, col2, rad2)
