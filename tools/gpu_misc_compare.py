"""fused-SSIM / simple-knn / sparse Adam vs the reference build on a GPU (diagnostic, prints JSON)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import refutil as U  # noqa: E402
import dogs_b200.fused_ssim as my_ssim  # noqa: E402
from dogs_b200.simple_knn._C import distCUDA2  # noqa: E402
import dogs_b200.diff_gaussian_rasterization as my_dgr  # noqa: E402


def main():
    dev = "cuda:0"
    dgr, rssim, rknn = U.ref_modules()
    rep = {}
    g = torch.Generator(device="cpu").manual_seed(0)
    for (H, W) in ((840, 1297), (67, 45)):
        a = torch.rand(1, 3, H, W, generator=g).to(dev)
        b = torch.rand(1, 3, H, W, generator=g).to(dev)
        for pad in ("same", "valid"):
            a1 = a.clone().requires_grad_(True)
            a2 = a.clone().requires_grad_(True)
            l1 = rssim.fused_ssim(a1, b, padding=pad)
            l2 = my_ssim.fused_ssim(a2, b, padding=pad)
            l1.backward(); l2.backward()
            rep[f"ssim_{H}x{W}_{pad}_val_absdiff"] = abs(l1.item() - l2.item())
            rep[f"ssim_{H}x{W}_{pad}_grad_maxabs"] = float((a1.grad - a2.grad).abs().max())
            rep[f"ssim_{H}x{W}_{pad}_grad_bits_mismatch"] = int((U.bits(a1.grad) != U.bits(a2.grad)).sum())
        m1 = rssim.fused_ssim(a, b, train=False)
        m2 = my_ssim.fused_ssim(a, b, train=False)
        rep[f"ssim_{H}x{W}_notrain_absdiff"] = abs(m1.item() - m2.item())
    for P in (100_000, 1000, 5):
        pts = torch.randn(P, 3, generator=g).to(dev)
        d1 = rknn.distCUDA2(pts)
        d2 = distCUDA2(pts)
        torch.cuda.synchronize()
        rep[f"knn_{P}_bits_mismatch"] = int((U.bits(d1) != U.bits(d2)).sum())
        rep[f"knn_{P}_maxrel"] = float(((d1 - d2).abs() / d1.abs().clamp_min(1e-30)).max())
    for (N, M) in ((10000, 45), (10000, 3), (777, 1), (1001, 4)):
        p = torch.randn(N, M, generator=g).to(dev)
        gr = torch.randn(N, M, generator=g).to(dev)
        m = torch.randn(N, M, generator=g).to(dev) * 0.1
        v = torch.rand(N, M, generator=g).to(dev) * 0.01
        vis = (torch.rand(N, generator=g) > 0.4).to(dev)
        pa, ma, va = p.clone(), m.clone(), v.clone()
        pb, mb, vb = p.clone(), m.clone(), v.clone()
        dgr._C.adamUpdate(pa, gr, ma, va, vis, 1e-3, 0.9, 0.999, 1e-15, N, M)
        my_dgr.adamUpdate(pb, gr, mb, vb, vis, 1e-3, 0.9, 0.999, 1e-15, N, M)
        torch.cuda.synchronize()
        rep[f"adam_{N}x{M}_bits_mismatch"] = int((U.bits(pa) != U.bits(pb)).sum() + (U.bits(ma) != U.bits(mb)).sum() + (U.bits(va) != U.bits(vb)).sum())
    print(json.dumps(rep, indent=1))
    with open(os.path.join(ROOT, "gpurun_out", "misc_compare.json"), "w") as f:
        json.dump(rep, f, indent=1)


if __name__ == "__main__":
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    main()
