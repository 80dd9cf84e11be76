"""Host-side cost of one training step (tiny scene => GPU time negligible): wall clock per step."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dogs_b200 import scenes
import dogs_b200.diff_gaussian_rasterization as dgr
from dogs_b200.fused_ssim import fused_ssim
dev = "cuda:0"
sc = scenes.make_scene("tiny", 0, dev)
target = torch.rand(3, sc["H"], sc["W"], device=dev)
params = {k: sc[k].clone().requires_grad_(True) for k in ("means3D", "dc", "sh", "opacities", "scales", "rotations")}
def step(with_loss=True):
    for v in params.values():
        v.grad = None
    img, radii, depth, spts = scenes.render(dgr, sc, params)
    if with_loss:
        loss = 0.8 * (img - target).abs().mean() + 0.2 * (1.0 - fused_ssim(img[None], target[None]))
    else:
        loss = img.sum()
    loss.backward()
for name, wl in (("fwd+loss+bwd", True), ("fwd+sum+bwd", False)):
    for _ in range(20): step(wl)
    torch.cuda.synchronize()
    ts = []
    for _ in range(200):
        t0 = time.perf_counter(); step(wl); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    ts.sort()
    print(name, "host+gpu wall per step: median %.3f ms, p10 %.3f, p90 %.3f" % (ts[100]*1e3, ts[20]*1e3, ts[180]*1e3))
with torch.no_grad():
    for _ in range(20): scenes.render(dgr, sc, params)
    torch.cuda.synchronize(); ts=[]
    for _ in range(200):
        t0 = time.perf_counter(); scenes.render(dgr, sc, params); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    ts.sort(); print("fwd only wall: median %.3f ms p10 %.3f p90 %.3f" % (ts[100]*1e3, ts[20]*1e3, ts[180]*1e3))
print("cpu count", os.cpu_count(), "load", os.getloadavg())
