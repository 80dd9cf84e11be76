"""Generates tests/golden/*.npz by running the UNMODIFIED reference CUDA build (oracle/_ref) on a
B200:  gpurun -- python tools/make_golden.py   (then copy gpurun_out/golden/*.npz to tests/golden/).

The reference ships no golden vectors of its own, so these fixtures are what pins the
CPU oracle (tests/test_oracle_golden.py) in the GPU-less container.  Inputs are stored next to the
outputs so the fixtures do not depend on torch's CPU RNG.  Gradients are taken with the reference's
max_contrib array set to the intended per-tile max of n_contrib (see
tests/test_gpu_rasterizer_parity.py for why) and, separately, to the zero-stale value."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import refutil as U  # noqa: E402
from dogs_b200 import scenes  # noqa: E402

DEV = "cuda:0"
OUT = os.path.join(ROOT, "gpurun_out", "golden")
os.makedirs(OUT, exist_ok=True)
dgr, rssim, rknn = U.ref_modules()


def npy(t):
    return t.detach().cpu().numpy()


def raster_case(name, cfg, seed, P, aa, bg):
    sc = scenes.make_scene(cfg, seed, DEV, P=P)
    H, W = sc["H"], sc["W"]
    N = H * W
    Pn = sc["means3D"].shape[0]
    bgt = torch.tensor(bg, dtype=torch.float32, device=DEV)
    r = U.ref_forward_raw(dgr, sc, bg=bgt, antialiasing=aa)
    torch.cuda.synchronize()
    rg = U.parse_ref_geom(r["geom"], Pn)
    rb = U.parse_ref_binning(r["binning"], r["num_rendered"])
    ri = U.parse_ref_image(r["img"], N)
    T = ((W + 15) // 16) * ((H + 15) // 16)
    keys = rb["keys"]
    valid = ((keys >> 32) & 0xFFFFFFFF) != 0xFFFFFFFF
    g = torch.Generator(device="cpu").manual_seed(1)
    dLc = torch.randn(3, H, W, generator=g).to(DEV)
    dLd = torch.randn(1, H, W, generator=g).to(DEV)
    out = dict(
        cfg=cfg, seed=seed, aa=int(aa), bg=np.asarray(bg, np.float32), W=W, H=H,
        tanfovx=sc["tanfovx"], tanfovy=sc["tanfovy"], sh_degree=sc["sh_degree"],
        **{"in_" + k: npy(sc[k]) for k in ("means3D", "dc", "sh", "opacities", "scales", "rotations",
                                            "viewmatrix", "projmatrix", "campos")},
        radii=npy(r["radii"]), means2D=npy(rg["means2D"]).reshape(Pn, 2), depths=npy(rg["depths"]),
        conic_opacity=npy(rg["conic_opacity"]).reshape(Pn, 4), rgb=npy(rg["rgb"]).reshape(Pn, 3),
        clamped=npy(rg["clamped"]).reshape(Pn, 3), cov3D=npy(rg["cov3D"]).reshape(Pn, 6),
        tiles_touched=npy(rg["tiles_touched"]), num_rendered_with_pads=r["num_rendered"],
        keys=npy(keys[valid]), point_list=npy(rb["point_list"][valid]),
        ranges=npy(ri["ranges"]).reshape(N, 2)[:T], n_contrib=npy(ri["n_contrib"]),
        final_T=npy(ri["accum_alpha"]), color=npy(r["color"]), invdepth=npy(r["invdepth"]),
        dL_dcolor=npy(dLc), dL_dinvdepth=npy(dLd))
    for mode, rows in (("intended", 16), ("zero_stale", 2)):
        ri["max_contrib"][:T] = U.tile_max(ri["n_contrib"], H, W, rows=rows)
        gr = U.ref_backward_raw(dgr, sc, r, dLc, dLd, bg=bgt, antialiasing=aa)
        torch.cuda.synchronize()
        for k, v in gr.items():
            out[f"grad_{mode}_{k}"] = npy(v)
    rast = dgr.GaussianRasterizer(scenes.raster_settings(dgr, sc))
    out["mark_visible"] = npy(rast.markVisible(sc["means3D"]))
    out["visible_filter"] = npy(rast.visible_filter(sc["means3D"], sc["scales"], sc["rotations"]))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "P", Pn, "R_valid", int(valid.sum()), "visible", int((r["radii"] > 0).sum()))


raster_case("raster_tiny_s0", "tiny", 0, 1200, False, (0.0, 0.0, 0.0))
raster_case("raster_tiny_s1_aa_bg", "tiny", 1, 800, True, (0.2, 0.5, 0.7))

# fused-SSIM
g = torch.Generator(device="cpu").manual_seed(3)
a = torch.rand(1, 3, 45, 70, generator=g).to(DEV)
b = (a.cpu() * 0.7 + 0.3 * torch.rand(1, 3, 45, 70, generator=g)).to(DEV)
fusedssim, fusedssim_backward = rssim.fusedssim, rssim.fusedssim_backward  # the reference's pybind entry points
m, d1, d2, d3 = fusedssim(0.01 ** 2, 0.03 ** 2, a, b, True)
dmap = torch.randn(1, 3, 45, 70, generator=g).to(DEV)
gimg = fusedssim_backward(0.01 ** 2, 0.03 ** 2, a, b, dmap, d1, d2, d3)
np.savez_compressed(os.path.join(OUT, "ssim.npz"), img1=npy(a), img2=npy(b), ssim_map=npy(m), dm_dmu1=npy(d1),
                    dm_dsigma1_sq=npy(d2), dm_dsigma12=npy(d3), dL_dmap=npy(dmap), dL_dimg1=npy(gimg),
                    mean_same=rssim.fused_ssim(a, b).item(), mean_valid=rssim.fused_ssim(a, b, padding="valid").item())

# simple-knn
pts = torch.randn(3000, 3, generator=g).to(DEV)
np.savez_compressed(os.path.join(OUT, "knn.npz"), points=npy(pts), dist2=npy(rknn.distCUDA2(pts)))

# sparse Adam
N, M = 500, 45
p = torch.randn(N, M, generator=g).to(DEV)
gr = torch.randn(N, M, generator=g).to(DEV)
m1 = (torch.randn(N, M, generator=g) * 0.1).to(DEV)
v1 = (torch.rand(N, M, generator=g) * 0.01).to(DEV)
vis = (torch.rand(N, generator=g) > 0.4).to(DEV)
po, mo, vo = p.clone(), m1.clone(), v1.clone()
dgr._C.adamUpdate(po, gr, mo, vo, vis, 1e-3, 0.9, 0.999, 1e-15, N, M)
torch.cuda.synchronize()
np.savez_compressed(os.path.join(OUT, "adam.npz"), param=npy(p), grad=npy(gr), exp_avg=npy(m1), exp_avg_sq=npy(v1),
                    visible=npy(vis), lr=1e-3, b1=0.9, b2=0.999, eps=1e-15, param_out=npy(po), exp_avg_out=npy(mo),
                    exp_avg_sq_out=npy(vo))
print("golden written to", OUT)
