"""Stage-by-stage comparison of this package against the reference build (oracle/_ref) on a GPU.
Diagnostic tool (prints counts, never asserts): python tools/gpu_stage_compare.py C1 [seed]"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import refutil as U  # noqa: E402
from dogs_b200 import scenes  # noqa: E402
import dogs_b200.diff_gaussian_rasterization as ours_mod  # noqa: E402


def main():
    cfg = sys.argv[1] if len(sys.argv) > 1 else "C1"
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    aa = len(sys.argv) > 3 and sys.argv[3] == "aa"
    dev = "cuda:0"
    sc = scenes.make_scene(cfg, seed, dev)
    P, H, W = sc["means3D"].shape[0], sc["H"], sc["W"]
    N = H * W
    rep = {"cfg": cfg, "seed": seed, "P": P, "aa": aa}
    dgr, _, _ = U.ref_modules()
    r = U.ref_forward_raw(dgr, sc, antialiasing=aa)
    torch.cuda.synchronize()
    o = U.ours_forward_raw(sc, antialiasing=aa)
    rg = U.parse_ref_geom(r["geom"], P)
    R_ref = r["num_rendered"]
    rb = U.parse_ref_binning(r["binning"], R_ref)
    ri = U.parse_ref_image(r["img"], N)
    vis = r["radii"] > 0
    rep["visible"] = int(vis.sum())
    rep["radii_mismatch"] = int((r["radii"] != o["radii"]).sum())
    m2 = rg["means2D"].view(P, 2)
    co = rg["conic_opacity"].view(P, 4)
    rgb = rg["rgb"].view(P, 3)
    rep["means2D_bits_mismatch"] = int((U.bits(m2)[vis] != U.bits(o["means2D"])[vis]).any(1).sum())
    rep["conic_bits_mismatch"] = int((U.bits(co)[vis] != U.bits(o["conic_opacity"])[vis]).any(1).sum())
    rep["depth_bits_mismatch"] = int((U.bits(rg["depths"])[vis] != U.bits(o["depths"])[vis]).sum())
    rep["rgb_bits_mismatch"] = int((U.bits(rgb)[vis] != U.bits(o["rgb"])[vis]).any(1).sum())
    rep["rgb_maxabs"] = float((rgb[vis] - o["rgb"][vis]).abs().max())
    cl_ref = rg["clamped"].view(P, 3)
    cl = (cl_ref[:, 0].int() | (cl_ref[:, 1].int() << 1) | (cl_ref[:, 2].int() << 2))
    rep["clamped_mismatch"] = int((cl[vis] != o["clamped"].int()[vis]).sum())
    # keys: strip reference pads
    keys = rb["keys"]
    valid = ((keys >> 32) & 0xFFFFFFFF) != 0xFFFFFFFF
    keys_v = keys[valid]
    pl_v = rb["point_list"][valid]
    rep["R_ref_with_pads"] = int(R_ref)
    rep["R_ref_valid"] = int(valid.sum())
    rep["R_ours"] = int(o["num_rendered"])
    # per-Gaussian instance counts
    cnt_ref = torch.bincount(pl_v.long(), minlength=P)
    rep["tiles_count_mismatch"] = int((cnt_ref != o["tiles"].long()).sum())
    if keys_v.numel() == o["num_rendered"]:
        my_keys = (o["tile_keys"].long() << 32) | (U.bits(o["depths"])[o["point_list"].long()].long() & 0xFFFFFFFF)
        rep["sorted_keys_mismatch"] = int((my_keys != keys_v).sum())
        rep["point_list_mismatch"] = int((pl_v != o["point_list"]).sum())
    T = o["ranges"].shape[0]
    rr = ri["ranges"].view(N, 2)[:T]
    rep["ranges_mismatch"] = int((rr != o["ranges"]).any(1).sum())
    rep["n_contrib_mismatch"] = int((ri["n_contrib"] != o["n_contrib"]).sum())
    rep["final_T_bits_mismatch"] = int((U.bits(ri["accum_alpha"]) != U.bits(o["final_T"])).sum())
    rep["color_maxabs"] = float((r["color"] - o["color"]).abs().max())
    rep["invdepth_maxabs"] = float((r["invdepth"] - o["invdepth"]).abs().max())
    rep["sum_n_contrib"] = int(ri["n_contrib"].long().sum())
    rep["num_buckets_ref"] = int(r["num_buckets"])
    del r, o, rg, rb, ri
    torch.cuda.empty_cache()

    # ---- backward through the public API on both ----
    grads = {}
    torch.manual_seed(1)
    target = torch.rand(3, H, W, device=dev)
    wdepth = torch.rand(1, H, W, device=dev)
    for name, mod in (("ref", dgr), ("ours", ours_mod), ("ours_quirk", ours_mod)):
        ours_mod.REFERENCE_BACKWARD_QUIRK = name == "ours_quirk"
        p = {k: sc[k].clone().requires_grad_(True) for k in ("means3D", "dc", "sh", "opacities", "scales", "rotations")}
        img, radii, depth, spts = scenes.render(mod, sc, p, antialiasing=aa)
        loss = (img - target).abs().mean() + 0.1 * (depth * wdepth).mean()
        loss.backward()
        torch.cuda.synchronize()
        grads[name] = {k: v.grad.detach().clone() for k, v in p.items()}
        grads[name]["means2D"] = spts.grad.detach().clone()
    for k in grads["ref"]:
        a, b = grads["ours"][k], grads["ref"][k]
        rep["grad_rel_" + k] = U.rel_to_max(a, b)
        rep["gradq_rel_" + k] = U.rel_to_max(grads["ours_quirk"][k], b)
        rep["grad_nan_" + k] = int(torch.isnan(a).sum())
    print(json.dumps(rep, indent=1))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"stage_compare_{cfg}_{seed}{'_aa' if aa else ''}.json"), "w") as f:
        json.dump(rep, f, indent=1)


if __name__ == "__main__":
    main()
