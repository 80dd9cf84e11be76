"""Per-tensor error of both float32 backwards against the FP64 truth (oracle/truth_bwd.cu) at full size.
    python tools/gpu_truth_compare.py C3 [runs]"""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import refutil as U
from dogs_b200 import scenes
cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
DEV = "cuda:0"
keys = ("means3D", "means2D", "dc", "sh", "opacities", "scales", "rotations")
dgr = U.ref_modules()[0]
sc = scenes.make_scene(cfg, 0, DEV)
H, W = sc["H"], sc["W"]
r = U.ref_forward_raw(dgr, sc)
ri = U.parse_ref_image(r["img"], H * W)
T = ((W + 15) // 16) * ((H + 15) // 16)
ri["max_contrib"][:T] = U.tile_max(ri["n_contrib"], H, W)
g = torch.Generator(device="cpu").manual_seed(1)
dLc = torch.randn(3, H, W, generator=g).to(DEV); dLd = torch.randn(1, H, W, generator=g).to(DEV)
o = U.ours_forward_raw(sc)
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
t0.record()
truth = U.truth_backward(sc, o, dLc, dLd)
t1.record(); torch.cuda.synchronize()
out = {"cfg": cfg, "truth_ms": t0.elapsed_time(t1)}
refs = []
for _ in range(2):
    gr = U.ref_backward_raw(dgr, sc, r, dLc, dLd)
    torch.cuda.synchronize()
    refs.append({k: gr[k].clone() for k in keys})
go = U.ours_backward(sc, dLc, dLd)
for k in keys:
    t = truth[k]
    e_o = U.err_vs_truth(go[k], t)
    e_r = [U.err_vs_truth(x[k].reshape(go[k].shape), t) for x in refs]
    tol = 1e-4 * float(t.abs().max())
    n_o = int(((go[k].double() - t.reshape(go[k].shape)).abs() > tol).sum())
    n_r = int(((refs[0][k].reshape(go[k].shape).double() - t.reshape(go[k].shape)).abs() > tol).sum())
    out[k] = {"ours": e_o, "ref": e_r, "ours_vs_ref": U.rel_to_max(go[k], refs[0][k].reshape(go[k].shape)),
              "n_out_ours": n_o, "n_out_ref": n_r, "max_abs_truth": float(t.abs().max())}
    print(f"{k:10s} ours-vs-truth {e_o:.2e} (#>1e-4: {n_o})   ref-vs-truth {e_r[0]:.2e} {e_r[1]:.2e} (#>1e-4: {n_r})   ours-vs-ref {out[k]['ours_vs_ref']:.2e}")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"truth_{cfg}.json"), "w"), indent=1)
