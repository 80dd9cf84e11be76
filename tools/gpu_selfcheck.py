"""Runs every kernel of this package once on small inputs (used under compute-sanitizer)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dogs_b200 import scenes  # noqa: E402
import dogs_b200.diff_gaussian_rasterization as dgr  # noqa: E402
import dogs_b200.fused_ssim as ssim  # noqa: E402
from dogs_b200.simple_knn._C import distCUDA2  # noqa: E402

dev = "cuda:0"
for cfg, P, aa in (("tiny", None, False), ("tiny", 37, True), ("C1", 20000, False)):
    sc = scenes.make_scene(cfg, 0, dev, P=P)
    p = {k: sc[k].clone().requires_grad_(True) for k in ("means3D", "dc", "sh", "opacities", "scales", "rotations")}
    img, radii, depth, spts = scenes.render(dgr, sc, p, antialiasing=aa, depth_threshold=0.23)
    target = torch.rand_like(img)
    loss = 0.8 * (img - target).abs().mean() + 0.2 * (1 - ssim.fused_ssim(img[None], target[None])) + depth.mean()
    loss.backward()
    torch.cuda.synchronize()
    print(cfg, P, "loss", loss.item(), "visible", int((radii > 0).sum()), "grad finite",
          all(torch.isfinite(v.grad).all().item() for v in p.values()))
    rast = dgr.GaussianRasterizer(scenes.raster_settings(dgr, sc))
    print(" markVisible", int(rast.markVisible(sc["means3D"]).sum()),
          "filter", int((rast.visible_filter(sc["means3D"], sc["scales"], sc["rotations"]) > 0).sum()))
pts = torch.randn(5000, 3, device=dev)
print("knn", distCUDA2(pts).mean().item())
N, M = 1000, 45
prm = torch.randn(N, M, device=dev)
dgr.adamUpdate(prm, torch.randn(N, M, device=dev), torch.zeros(N, M, device=dev), torch.zeros(N, M, device=dev),
               torch.rand(N, device=dev) > 0.5, 1e-3, 0.9, 0.999, 1e-15, N, M)
torch.cuda.synchronize()
print("selfcheck done")
