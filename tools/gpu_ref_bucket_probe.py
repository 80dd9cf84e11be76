"""Probe the reference backward per tile-list entry: gradient w.r.t. colour restricted to one tile."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import refutil as U
from dogs_b200 import scenes
from oracle import oracle as O
dev = "cuda:0"
sc = scenes.make_scene("tiny", 0, dev)
H, W, P = sc["H"], sc["W"], sc["means3D"].shape[0]
dgr = U.ref_modules()[0]
r = U.ref_forward_raw(dgr, sc)
torch.cuda.synchronize()
s = {k: (v.cpu().numpy() if isinstance(v, torch.Tensor) else v) for k, v in sc.items()}
pre, bins, im = O.forward(s)
gx = W // 16
e = torch.Tensor([])
for tile in [int(a) for a in sys.argv[1:]] or [4, 30, 17]:
    ty, tx = divmod(tile, gx)
    g = torch.zeros(3, H, W, device=dev)
    torch.manual_seed(tile)
    g[:, ty * 16:ty * 16 + 16, tx * 16:tx * 16 + 16] = torch.randn(3, 16, 16, device=dev)
    gd = torch.zeros(1, H, W, device=dev)
    out = dgr._C.rasterize_gaussians_backward(
        torch.zeros(3, device=dev), sc["means3D"], r["radii"], e, sc["opacities"], sc["scales"], sc["rotations"], 1.0, e,
        sc["viewmatrix"], sc["projmatrix"], sc["tanfovx"], sc["tanfovy"], g, sc["dc"], sc["sh"], gd, 3, sc["campos"],
        r["geom"], r["num_rendered"], r["binning"], r["img"], r["num_buckets"], r["sample"], False, False)
    torch.cuda.synchronize()
    dL_dcol_ref = out[1].cpu().numpy()          # [P,3] raw sum_p weight * dL/dC
    dL_dop_ref = out[2].cpu().numpy()[:, 0]
    acc = O.backward(s, pre, bins, im, g.cpu().numpy(), None)
    # raw oracle accumulators: recompute
    import ctypes as C
    accd = np.zeros((P, 10), np.float64)
    O.lib().orc_render_bwd(C.c_int(W), C.c_int(H), O._p(bins["ranges"]), O._p(bins["point_list"]), O._p(np.zeros(3, np.float32)),
                           O._p(pre["means2D"]), O._p(pre["conic_opacity"]), O._p(pre["rgb"]), O._p(pre["depths"]),
                           O._p(im["final_T"]), O._p(im["n_contrib"]), O._p(im["color"]), O._p(im["invdepth"]),
                           O._p(np.ascontiguousarray(g.cpu().numpy())), None, O._p(accd), C.c_int(0), None)
    lo, hi = bins["ranges"][tile]
    nc = im["n_contrib"].reshape(H, W)[ty * 16:ty * 16 + 16, tx * 16:tx * 16 + 16]
    print(f"tile {tile} len {hi-lo} n_contrib min {nc.min()} max {nc.max()}")
    for k in range(lo, hi):
        gid = bins["point_list"][k]
        print(f"  entry {k-lo:3d} lane {(k-lo)%32:2d} gauss {gid:5d}  ref_col {dL_dcol_ref[gid]}  true_col {accd[gid,6:9].astype(np.float32)}  ref_op {dL_dop_ref[gid]:+.4e} true_op {accd[gid,5]:+.4e}")
