"""Golden vectors for densification / pruning / opacity reset, produced by EXECUTING the reference's own
GaussianSplatModel methods (unmodified, imported from /root/reference with tools/ref_import.py, CPU tensors):

  add_densification_stats                       gaussian_splat_model.py:529-541 (+ gaussian_trainer.py:433-437)
  densify_and_prune -> densify_and_clone, densify_and_split, densification_postfix, prune_points   :369-527
  cat_tensors_to_optimizer / prune_optimizer / replace_tensor_to_optimizer                         :34-108
  reset_opacity                                 :362-367
  prune_gaussians_with_opt + calculate_v_imp_score   :418-423, prune.py:14-31

The optimizer is a torch.optim.Adam with the reference's six named groups and non-trivial moments.  The
split samples are what torch.normal returns in that call after torch.manual_seed(SEED): the generator
re-draws them the same way and stores them as an input.

    python tools/make_golden_densify.py      ->  tests/golden/densify.npz
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ref_import  # noqa: E402

GaussianSplatModel = ref_import.reference_classes()[0]
import importlib  # noqa: E402
prune_mod = importlib.import_module("conerf.model.gaussian_fields.prune")

NAMES = ("xyz", "f_dc", "f_rest", "opacity", "scaling", "quaternion")
SEED = 777


def build(P, g):
    m = GaussianSplatModel(max_sh_degree=3, percent_dense=0.01, device="cpu")
    xyz = torch.randn(P, 3, generator=g)
    m.init_from_external_properties(
        xyz, torch.randn(P, 1, 3, generator=g), torch.randn(P, 15, 3, generator=g) * 0.1,
        # scaling: log-normal around percent_dense * extent so that both clone and split fire
        torch.randn(P, 3, generator=g) * 0.8 + np.log(0.006), torch.randn(P, 4, generator=g),
        torch.randn(P, 1, generator=g) * 2.5 - 1.0, optimizable=True)
    params = dict(xyz=m._xyz, f_dc=m._features_dc, f_rest=m._features_rest, opacity=m._opacity, scaling=m._scaling,
                  quaternion=m._quaternion)
    opt = torch.optim.Adam([{"params": [params[n]], "lr": 1e-3, "name": n} for n in NAMES], lr=0.0, eps=1e-15)
    for n in NAMES:
        opt.state[params[n]] = {"step": torch.tensor(5.0), "exp_avg": torch.randn(params[n].shape, generator=g),
                                "exp_avg_sq": torch.rand(params[n].shape, generator=g)}
    m.xyz_gradient_accum = torch.zeros(P, 1)
    m.denom = torch.zeros(P, 1)
    m.max_radii2D = torch.zeros(P)
    return m, opt


def dump(out, prefix, m, opt):
    for n, t in zip(NAMES, (m._xyz, m._features_dc, m._features_rest, m._opacity, m._scaling, m._quaternion)):
        out[f"{prefix}_{n}"] = t.detach().numpy().copy()
        st = opt.state[t]
        out[f"{prefix}_m_{n}"] = st["exp_avg"].numpy().copy()
        out[f"{prefix}_v_{n}"] = st["exp_avg_sq"].numpy().copy()
    out[f"{prefix}_accum"] = m.xyz_gradient_accum.numpy().copy()
    out[f"{prefix}_denom"] = m.denom.numpy().copy()
    out[f"{prefix}_max_radii2D"] = m.max_radii2D.numpy().copy()


def main():
    g = torch.Generator(device="cpu").manual_seed(11)
    P = 400
    out = {"P": P, "extent": 1.2, "max_grad": 0.0002, "min_opacity": 0.005, "max_screen_size": 20.0, "seed": SEED,
           "bbox_zmin": -1.8}
    m, opt = build(P, g)
    dump(out, "in", m, opt)
    # three iterations of statistics (gaussian_trainer.py:433-439)
    for it in range(3):
        radii = (torch.rand(P, generator=g) * 30).to(torch.int32)
        radii[torch.rand(P, generator=g) < 0.3] = 0
        grad2d = torch.randn(P, 3, generator=g) * 0.0004
        grad2d[::17] = 0.0
        out[f"stat{it}_radii"], out[f"stat{it}_grad2d"] = radii.numpy().copy(), grad2d.numpy().copy()
        vis = radii > 0
        m.max_radii2D[vis] = torch.max(m.max_radii2D[vis], radii[vis])
        pts = type("P", (), {"grad": grad2d})()
        m.add_densification_stats(pts, vis)
    out["stats_accum"], out["stats_denom"] = m.xyz_gradient_accum.numpy().copy(), m.denom.numpy().copy()
    out["stats_max_radii2D"] = m.max_radii2D.numpy().copy()

    # the samples densify_and_split will draw: same seed, same call, same shapes
    grads = m.xyz_gradient_accum / m.denom
    grads[grads.isnan()] = 0.0
    n_clone = int(((torch.norm(grads, dim=-1) >= out["max_grad"]) &
                   (m.get_scaling.max(dim=1).values <= m.percent_dense * out["extent"])).sum())
    sel = (grads.squeeze() >= out["max_grad"]) & (m.get_scaling.max(dim=1).values > m.percent_dense * out["extent"])
    stds = m.get_scaling[sel].detach().repeat(2, 1)
    torch.manual_seed(SEED)
    out["samples"] = torch.normal(mean=torch.zeros((stds.size(0), 3)), std=stds).numpy().copy()
    out["n_clone"], out["n_split"] = n_clone, int(sel.sum())

    for case, (mss, bbox) in {"a": (out["max_screen_size"], None), "b": (None, [0.0, 0.0, out["bbox_zmin"]])}.items():
        g2 = torch.Generator(device="cpu").manual_seed(11)
        m2, opt2 = build(P, g2)
        m2.xyz_gradient_accum, m2.denom = m.xyz_gradient_accum.clone(), m.denom.clone()
        m2.max_radii2D = m.max_radii2D.clone()
        torch.manual_seed(SEED)
        m2.densify_and_prune(max_grad=out["max_grad"], min_opacity=out["min_opacity"], extent=out["extent"],
                             max_screen_size=mss, optimizer=opt2, bounding_box=bbox)
        dump(out, f"dp_{case}", m2, opt2)
        print(f"case {case}: {P} -> {m2.get_xyz.shape[0]} Gaussians (clone {n_clone}, split {int(sel.sum())})")
        if case == "a":
            m2.reset_opacity(opt2)
            dump(out, "reset", m2, opt2)
            # LightGaussian prune (gaussian_trainer.py:456-469)
            n2 = m2.get_xyz.shape[0]
            imp = torch.rand(n2, generator=g) * 50
            out["imp_list"] = imp.numpy().copy()
            v_list = prune_mod.calculate_v_imp_score(m2, imp, 0.1)
            out["v_list"] = v_list.detach().numpy().copy()
            m2.prune_gaussians_with_opt(0.6 * 0.5, v_list, opt2)
            dump(out, "lg", m2, opt2)
            print(f"  LightGaussian prune: {n2} -> {m2.get_xyz.shape[0]}")
    path = os.path.join(ROOT, "tests", "golden", "densify.npz")
    np.savez_compressed(path, **out)
    print("written", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
