"""Where a training step's time goes, with and without the per-iteration ADMM work:
CUDA events between render / loss / backward / penalty, median over N steps.  python tools/gpu_step_segments.py [C3]"""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dogs_b200 import scenes, admm as A
import dogs_b200.diff_gaussian_rasterization as dgr
from dogs_b200.fused_ssim import fused_l1_ssim_loss
cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
dev = "cuda:0"
sc = scenes.make_scene(cfg, 0, dev)
P, H, W = sc["means3D"].shape[0], sc["H"], sc["W"]
target = torch.rand(3, H, W, device=dev)
leaves = {"means3D": sc["means3D"].clone(), "dc": sc["dc"].reshape(P, 3).clone(), "sh": sc["sh"].reshape(P, 45).clone(),
          "opacities": sc["opacities"].clone(), "scales": sc["scales"].clone(), "rotations": sc["rotations"].clone()}
for v in leaves.values():
    v.requires_grad_(True)
def views():
    d = dict(leaves); d["dc"] = leaves["dc"].view(P, 1, 3); d["sh"] = leaves["sh"].view(P, 15, 3); return d
def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e
def run(blk, n=30):
    segs, host = [], []
    for it in range(n + 5):
        t0 = time.perf_counter()
        e = [ev()]
        for v in leaves.values():
            v.grad = None
        img, radii, depth, spts = scenes.render(dgr, sc, views()); e.append(ev())
        loss = fused_l1_ssim_loss(img, target, 0.2); e.append(ev())
        loss.backward(); e.append(ev())
        if blk is not None:
            blk.add_penalty_grads_()
        e.append(ev())
        host.append(time.perf_counter() - t0)
        if it >= 5:
            segs.append(e)
    torch.cuda.synchronize()
    names = ["render", "loss", "backward", "penalty"]
    med = lambda xs: sorted(xs)[len(xs) // 2]
    out = {nm: med([s[i].elapsed_time(s[i + 1]) for s in segs]) for i, nm in enumerate(names)}
    out["step(start-to-start)"] = med([segs[i][0].elapsed_time(segs[i + 1][0]) for i in range(len(segs) - 1)])
    out["host enqueue per step"] = med(host[5:]) * 1e3
    return out
print("plain   ", {k: round(v, 3) for k, v in run(None).items()})
ov = int(0.245 * P)
count = torch.ones(P, dtype=torch.int32); count[:ov] = 2; count[-ov:] = 2
plan = A.ConsensusPlan(torch.arange(P, dtype=torch.int64), count).to(dev)
blk = A.AdmmBlock([leaves[k] for k in ("means3D", "dc", "sh", "scales", "rotations", "opacities")], plan)
for mode in ((True, True), (True, False)):
    blk.fuse_into_rasterizer(*mode)
    print("admm fused all_groups=%s" % mode[1], {k: round(v, 3) for k, v in run(blk).items()})
blk.fuse_into_rasterizer(False)
print("admm unfused", {k: round(v, 3) for k, v in run(blk).items()})
