"""Per-stage device times of the rasterizer forward and backward (library stage-timing hook, CUDA events):
    python tools/gpu_stages.py C3 [reps]"""
import ctypes as C, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import refutil as U
from dogs_b200 import scenes, _lib as L
from dogs_b200.diff_gaussian_rasterization import _params, GaussianRasterizationSettings
cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 15
DEV = "cuda:0"
sc = scenes.make_scene(cfg, 0, DEV)
P, H, W = sc["means3D"].shape[0], sc["H"], sc["W"]
N = H * W
lib = L.lib()
names_f = ["preprocess", "depth_sort", "gather_scan_sync", "emit", "tile_sort", "ranges_order", "blend_fwd"]
names_b = ["blend_bwd", "preprocess_bwd"]
fw = {n: [] for n in names_f}
bw = {n: [] for n in names_b}
bg = torch.zeros(3, device=DEV)
rs = GaussianRasterizationSettings(H, W, sc["tanfovx"], sc["tanfovy"], bg, 1.0, sc["viewmatrix"], sc["projmatrix"], 3, sc["campos"], False, False, False, 0.0)
dLc = torch.randn(3, H, W, device=DEV) / N
new = lambda *s: torch.empty(s, device=DEV)
outs = [new(P, 3), new(P, 3), new(P, 3), new(P, 45), None, new(P, 1), new(P, 3), new(P, 4), None]
lib.dogs_stage_timing(1)
buf = (C.c_float * 16)()
kb = 0
for it in range(reps + 3):
    o = U.ours_forward_raw(sc, depth_key_bits=kb)      # exact (synchronising) path so that the stages line up
    kb = o["depth_key_bits"]
    n = lib.dogs_forward_stage_times(buf, 16)
    if it >= 3:
        for i in range(n):
            fw[names_f[i]].append(buf[i])
    al, keep = o["_keep"]
    prm, keep2 = _params(rs, P, 15, DEV)
    scratch = L.Allocator(DEV)
    L.check(lib.dgr_backward(C.byref(prm), o["num_rendered"], sc["means3D"].data_ptr(), sc["dc"].data_ptr(), sc["sh"].data_ptr(), None,
                             sc["opacities"].data_ptr(), sc["scales"].data_ptr(), sc["rotations"].data_ptr(), None,
                             o["radii"].data_ptr(), al[0].last.data_ptr(), al[1].last.data_ptr(), al[2].last.data_ptr(),
                             dLc.data_ptr(), None, *[L.ptr(t) for t in outs], scratch.cb, None, L.stream_ptr(DEV)))
    L.check(lib.dogs_stage_times(buf, 2))
    if it >= 3:
        bw["blend_bwd"].append(buf[0]); bw["preprocess_bwd"].append(buf[1])
    del o
lib.dogs_stage_timing(0)
med = lambda v: sorted(v)[len(v) // 2] if v else None
rep = {"cfg": cfg, "depth_key_bits": kb}
tot = 0.0
for k in names_f + names_b:
    m = med(fw.get(k) or bw.get(k))
    rep[k] = m
    if m:
        tot += m
rep["sum_ms"] = tot
print(json.dumps(rep))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", f"stages_{cfg}.json"), "w") as f:
    json.dump(rep, f)
