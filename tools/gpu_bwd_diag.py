"""Triangulates backward differences: reference CUDA vs ours vs CPU oracle on one scene."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import refutil as U
from dogs_b200 import scenes
import dogs_b200.diff_gaussian_rasterization as ours_mod
from oracle import oracle as O

cfg = sys.argv[1] if len(sys.argv) > 1 else "tiny"
use_depth = (sys.argv[2] != "nodepth") if len(sys.argv) > 2 else True
dev = "cuda:0"
sc = scenes.make_scene(cfg, 0, dev)
H, W, P = sc["H"], sc["W"], sc["means3D"].shape[0]
torch.manual_seed(1)
target = torch.rand(3, H, W, device=dev)
wdepth = torch.rand(1, H, W, device=dev)
dgr = U.ref_modules()[0]
grads = {}
gouts = {}
for name, mod in (("ref", dgr), ("ours", ours_mod)):
    p = {k: sc[k].clone().requires_grad_(True) for k in ("means3D", "dc", "sh", "opacities", "scales", "rotations")}
    img, radii, depth, spts = scenes.render(mod, sc, p)
    img.retain_grad(); depth.retain_grad()
    loss = (img - target).abs().mean() + (0.1 * (depth * wdepth).mean() if use_depth else 0)
    loss.backward()
    torch.cuda.synchronize()
    grads[name] = {k: v.grad.detach().cpu().numpy() for k, v in p.items()}
    grads[name]["means2D"] = spts.grad.detach().cpu().numpy()
    gouts[name] = (img.grad.detach().cpu().numpy(), depth.grad.detach().cpu().numpy() if depth.grad is not None else None, radii.cpu().numpy())
s = {k: (v.cpu().numpy() if isinstance(v, torch.Tensor) else v) for k, v in sc.items()}
pre, bins, im = O.forward(s)
go = O.backward(s, pre, bins, im, gouts["ref"][0], gouts["ref"][1])
grads["orc"] = go
def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
for k in ("means2D", "dc", "sh", "opacities", "means3D", "scales", "rotations"):
    print(f"{k:10s} ours-vs-ref {rel(grads['ours'][k], grads['ref'][k]):.3e}  orc-vs-ref {rel(grads['orc'][k].reshape(grads['ref'][k].shape), grads['ref'][k]):.3e}  ours-vs-orc {rel(grads['ours'][k], grads['orc'][k].reshape(grads['ours'][k].shape)):.3e}")
d = np.abs(grads["ours"]["means2D"] - grads["ref"]["means2D"]).max(1)
worst = np.argsort(-d)[:8]
tiles = bins["counts"]
for i in worst:
    print("idx", i, "diff", d[i], "ref", grads["ref"]["means2D"][i], "ours", grads["ours"]["means2D"][i], "orc", go["means2D"][i],
          "opacity", s["opacities"][i, 0], "radius", pre["radii"][i], "tiles", tiles[i], "xy", pre["means2D"][i], "depth", pre["depths"][i])
print("n wrong (>1e-3 of max):", int((d > 1e-3 * np.abs(grads['ref']['means2D']).max()).sum()), "of", P)
