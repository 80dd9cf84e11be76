"""dogs_b200 — B200-native (sm_100a) implementation of the DOGS rasterizer hot path.

Sub-packages mirror the reference's extension modules one to one:
  dogs_b200.diff_gaussian_rasterization  <- submodules/diff-gaussian-rasterization
  dogs_b200.fused_ssim                   <- submodules/fused-ssim
  dogs_b200.simple_knn                   <- submodules/simple-knn   (simple_knn._C.distCUDA2)
  dogs_b200.admm                         <- conerf/trainers/{slave,master}_gaussian_trainer.py (ADMM step)
`dogs_b200.install_dropin()` registers them under the reference's import names.
"""
import sys

__version__ = "0.1.0"


def install_dropin():
    """Makes `import diff_gaussian_rasterization`, `import fused_ssim` and
    `from simple_knn._C import distCUDA2` resolve to this package (what conerf imports:
    conerf/render/gaussian_render.py:8-11, conerf/trainers/gaussian_trainer.py:26,
    conerf/model/gaussian_fields/gaussian_splat_model.py:12)."""
    from . import diff_gaussian_rasterization, fused_ssim, simple_knn
    from .simple_knn import _C
    sys.modules["diff_gaussian_rasterization"] = diff_gaussian_rasterization
    sys.modules["fused_ssim"] = fused_ssim
    sys.modules["simple_knn"] = simple_knn
    sys.modules["simple_knn._C"] = _C
