"""How repeatable is the reference's own backward at full size?  ref-vs-ref (atomic order) next to ours-vs-ref."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import refutil as U
from dogs_b200 import scenes
cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
DEV = "cuda:0"
dgr = U.ref_modules()[0]
sc = scenes.make_scene(cfg, 0, DEV)
H, W = sc["H"], sc["W"]
r = U.ref_forward_raw(dgr, sc)
ri = U.parse_ref_image(r["img"], H * W)
T = ((W + 15) // 16) * ((H + 15) // 16)
ri["max_contrib"][:T] = U.tile_max(ri["n_contrib"], H, W)
g = torch.Generator(device="cpu").manual_seed(1)
dLc = torch.randn(3, H, W, generator=g).to(DEV); dLd = torch.randn(1, H, W, generator=g).to(DEV)
keys = ("means3D", "means2D", "dc", "sh", "opacities", "scales", "rotations")
runs = []
for _ in range(3):
    gr = U.ref_backward_raw(dgr, sc, r, dLc, dLd)
    torch.cuda.synchronize()
    runs.append({k: gr[k].clone() for k in keys})
go = U.ours_backward(sc, dLc, dLd)
go2 = U.ours_backward(sc, dLc, dLd)
for k in keys:
    a = runs[0][k]
    rr = max(U.rel_to_max(runs[1][k], a), U.rel_to_max(runs[2][k], a))
    print("outliers>1e-4: ref-vs-ref", U.count_outliers(runs[1][k], a, 1e-4), U.count_outliers(runs[2][k], a, 1e-4), "ours-vs-ref", U.count_outliers(go[k], a.reshape(go[k].shape), 1e-4))
    print(f"{k:10s} ref-vs-ref {rr:.2e}   ours-vs-ref {U.rel_to_max(go[k], a.reshape(go[k].shape)):.2e}   ours-vs-ours {U.rel_to_max(go2[k], go[k]):.2e}   max|g| {a.abs().max().item():.3e}")
    d = (go[k] - a.reshape(go[k].shape)).abs()
    i = int(d.reshape(d.shape[0], -1).max(1).values.argmax())
    print("    worst gaussian", i, "ours", go[k][i].flatten()[:4].tolist(), "ref", a.reshape(go[k].shape)[i].flatten()[:4].tolist(), "ref2", runs[1][k].reshape(go[k].shape)[i].flatten()[:4].tolist())
