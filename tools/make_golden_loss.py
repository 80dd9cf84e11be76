"""Golden vectors for the two per-iteration fusions, produced by the REFERENCE composition on a B200:
  loss_l1_ssim.npz  (1-l)*l1_loss + l*(1 - fused_ssim) with the reference's fused_ssim (gaussian_trainer.py:403-405)
  activations.npz   torch.sigmoid / torch.exp / F.normalize of raw parameters (gaussian_splat_model.py:148-153)
Run where oracle/_ref exists:  python tools/make_golden_loss.py   (writes tests/golden/)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import refutil as U  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
DEV = "cuda:0"
npy = lambda t: t.detach().cpu().numpy()
_, rssim, _ = U.ref_modules()

g = torch.Generator(device="cpu").manual_seed(17)
img = torch.rand(3, 53, 78, generator=g).to(DEV).requires_grad_(True)
gt = (img.detach().cpu() * 0.6 + 0.4 * torch.rand(3, 53, 78, generator=g)).to(DEV)
lam = 0.2
loss = (1.0 - lam) * (img - gt).abs().mean() + lam * (1.0 - rssim.fused_ssim(img[None], gt[None]))
loss.backward()
np.savez_compressed(os.path.join(OUT, "loss_l1_ssim.npz"), img=npy(img), gt=npy(gt), lam=lam, loss=loss.item(),
                    grad=npy(img.grad))

P = 4000
op = (torch.randn(P, 1, generator=g) * 3).to(DEV)
sc = (torch.randn(P, 3, generator=g) * 1.5 - 4).to(DEV)
rot = (torch.randn(P, 4, generator=g) * torch.rand(P, 1, generator=g) * 3).to(DEV)
rot[:2] = 0.0
np.savez_compressed(os.path.join(OUT, "activations.npz"), raw_opacity=npy(op), raw_scaling=npy(sc), raw_quaternion=npy(rot),
                    opacity=npy(torch.sigmoid(op)), scaling=npy(torch.exp(sc)),
                    quaternion=npy(torch.nn.functional.normalize(rot)))
print("written", OUT)
