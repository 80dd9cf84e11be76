"""Is the reference's backward self-consistent?  Finite differences on selected Gaussians (tiny scene)."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import refutil as U
from dogs_b200 import scenes
import dogs_b200.diff_gaussian_rasterization as ours_mod
dev = "cuda:0"
sc = scenes.make_scene("tiny", 0, dev)
H, W, P = sc["H"], sc["W"], sc["means3D"].shape[0]
torch.manual_seed(1)
target = torch.rand(3, H, W, device=dev)
wimg = torch.randn(3, H, W, device=dev)   # linear loss => finite differences are exact up to fp
dgr = U.ref_modules()[0]
def run(mod, sc2):
    p = {k: sc2[k].clone().requires_grad_(True) for k in ("means3D", "dc", "sh", "opacities", "scales", "rotations")}
    img, radii, depth, spts = scenes.render(mod, sc2, p)
    loss = (img * wimg).sum()
    loss.backward()
    torch.cuda.synchronize()
    return loss.item(), {k: v.grad.detach().clone() for k, v in p.items()}, img.detach().double()
l0, g_ref, img0 = run(dgr, sc)
_, g_ref2, _ = run(dgr, sc)
_, g_ours, _ = run(ours_mod, sc)
print("ref deterministic:", {k: float((g_ref[k] - g_ref2[k]).abs().max()) for k in g_ref})
for i in (811, 1173, 1760, 169, 5, 100, 1000):
    for ch in range(3):
        sc2 = dict(sc); dc = sc["dc"].clone(); eps = 1e-2; dc[i, 0, ch] += eps; sc2["dc"] = dc
        with torch.no_grad():
            img1 = scenes.render(dgr, sc2)[0].double()
        fd = ((img1 - img0) * wimg.double()).sum().item() / eps
        print(f"gauss {i} ch{ch}: FD {fd:+.6e}  ref {g_ref['dc'][i,0,ch].item():+.6e}  ours {g_ours['dc'][i,0,ch].item():+.6e}")
