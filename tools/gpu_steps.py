"""A few fwd + loss + bwd steps of one config through the public API (profiling target):
    python tools/gpu_steps.py C3 [steps]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dogs_b200 import scenes
import dogs_b200.diff_gaussian_rasterization as dgr
from dogs_b200.fused_ssim import fused_l1_ssim_loss
cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
dev = "cuda:0"
sc = scenes.make_scene(cfg, 0, dev)
torch.manual_seed(1234)
target = torch.rand(3, sc["H"], sc["W"], device=dev)
p = {k: sc[k].clone().requires_grad_(True) for k in ("means3D", "dc", "sh", "opacities", "scales", "rotations")}
for it in range(steps):
    for v in p.values():
        v.grad = None
    img, radii, depth, pts = scenes.render(dgr, sc, p)
    loss = fused_l1_ssim_loss(img, target, 0.2)
    loss.backward()
torch.cuda.synchronize()
print("ok", float(loss))
