"""Densification, pruning and opacity reset of a Gaussian model on the GPU.

Host-side mirror of the reference's GaussianSplatModel methods
(conerf/model/gaussian_fields/gaussian_splat_model.py:362-541) and of the optimizer surgery they call
(:34-108), over a plain SoA instead of nn.Parameters spread over optimizer groups:

    model = GaussianSoA(xyz, f_dc, f_rest, opacity, scaling, quaternion)      # raw (pre-activation) tensors
    model.add_densification_stats(grad_means2D, radii)                        # every iteration   (:529-541)
    model.densify_and_prune(max_grad, min_opacity, extent, max_screen_size)   # every 100 iters   (:499-527)
    model.reset_opacity()                                                     # every 3000 iters  (:362-367)
    model.prune_gaussians(percent, import_score)                              # LightGaussian     (:418-431)

densify_and_prune is ONE pass (classify -> plan -> scan -> gather, csrc/densify.cu) where the reference
concatenates, masks and re-concatenates the 6 parameter and 12 Adam-state tensors four times; the resulting
rows, their order and the Adam moments are those of the reference.  The split samples are drawn exactly as
the reference draws them — torch.normal(mean=0, std=scaling[selected].repeat(2, 1)) on the model's device —
so a run with the same seed consumes the same random stream.
"""
import ctypes as C
from typing import Dict, Optional

import torch

from . import _lib as L
from .compaction import compact

PARAMS = ("xyz", "f_dc", "f_rest", "opacity", "scaling", "quaternion")      # the reference's optimizer group names


def calculate_v_imp_score(scaling_activated: torch.Tensor, imp_list: torch.Tensor, v_pow: float) -> torch.Tensor:
    """prune.py:14-31 (LightGaussian volume-weighted importance); a handful of torch ops, once per prune."""
    volume = torch.prod(scaling_activated, dim=1)
    index = int(len(volume) * 0.9)
    sorted_volume, _ = torch.sort(volume, descending=True)
    kth_percent_largest = sorted_volume[index]
    return torch.pow(volume / kth_percent_largest, v_pow) * imp_list


class GaussianSoA:
    """The six raw parameter tensors of a 3DGS model, their Adam moments and the densification statistics."""

    def __init__(self, xyz, f_dc, f_rest, opacity, scaling, quaternion, percent_dense: float = 0.01,
                 with_adam_state: bool = True):
        self.p: Dict[str, torch.Tensor] = dict(xyz=xyz, f_dc=f_dc, f_rest=f_rest, opacity=opacity, scaling=scaling,
                                               quaternion=quaternion)
        for k, t in self.p.items():
            self.p[k] = L.require_cuda(t.detach(), k)
        P = self.p["xyz"].shape[0]
        for k, t in self.p.items():
            if t.shape[0] != P:
                raise RuntimeError(f"{k} has {t.shape[0]} rows, xyz has {P}")
        dev = self.p["xyz"].device
        self.percent_dense = percent_dense
        self.exp_avg = {k: torch.zeros_like(t) for k, t in self.p.items()} if with_adam_state else None
        self.exp_avg_sq = {k: torch.zeros_like(t) for k, t in self.p.items()} if with_adam_state else None
        self.xyz_gradient_accum = torch.zeros(P, 1, device=dev)
        self.denom = torch.zeros(P, 1, device=dev)
        self.max_radii2D = torch.zeros(P, device=dev)

    @property
    def num_gaussians(self):
        return self.p["xyz"].shape[0]

    @property
    def device(self):
        return self.p["xyz"].device

    def _every_tensor(self):
        out = [self.p[k] for k in PARAMS]
        if self.exp_avg is not None:
            out += [self.exp_avg[k] for k in PARAMS] + [self.exp_avg_sq[k] for k in PARAMS]
        return out

    def _set_every_tensor(self, ts):
        n = len(PARAMS)
        self.p = dict(zip(PARAMS, ts[:n]))
        if self.exp_avg is not None:
            self.exp_avg = dict(zip(PARAMS, ts[n:2 * n]))
            self.exp_avg_sq = dict(zip(PARAMS, ts[2 * n:3 * n]))

    # ---- every iteration ---------------------------------------------------------------------------------------
    @torch.no_grad()
    def add_densification_stats(self, grad_means2D: torch.Tensor, radii: torch.Tensor):
        """gaussian_trainer.py:433-439 + add_densification_stats (:529-541) in one launch: for the visible
        Gaussians (radii > 0) accum += |grad_means2D[:, :2]|, denom += 1, max_radii2D = max(., radii)."""
        P = self.num_gaussians
        g = L.require_cuda(grad_means2D, "grad_means2D")
        if g.shape != (P, 3) or radii.shape != (P,) or radii.dtype != torch.int32:
            raise RuntimeError("add_densification_stats: grad_means2D [P,3] float32 and radii [P] int32 expected")
        with torch.cuda.device(self.device):
            L.check(L.lib().dgr_densification_stats(P, g.data_ptr(), radii.contiguous().data_ptr(),
                                                    self.xyz_gradient_accum.data_ptr(), self.denom.data_ptr(),
                                                    self.max_radii2D.data_ptr(), L.stream_ptr(self.device)))

    # ---- every densification_interval iterations --------------------------------------------------------------
    @torch.no_grad()
    def densify_and_prune(self, max_grad: float, min_opacity: float, extent: float, max_screen_size: Optional[float],
                          bounding_box=None, num_replica: int = 2, generator: Optional[torch.Generator] = None,
                          samples: Optional[torch.Tensor] = None):
        """densify_and_prune (:499-527) = densify_and_clone + densify_and_split + the final prune.
        `samples` ([num_replica * n_split, 3], optional) replaces the torch.normal draw (tests, replay)."""
        P, dev = self.num_gaussians, self.device
        if P == 0:
            return 0
        lib = L.lib()
        cls = torch.empty(P, dtype=torch.uint8, device=dev)
        split_rank = torch.empty(P, dtype=torch.int32, device=dev)
        n_split = C.c_int64(0)
        scratch = L.Allocator(dev)
        p = self.p
        with torch.cuda.device(dev):
            L.check(lib.dgr_densify_classify(P, self.xyz_gradient_accum.data_ptr(), self.denom.data_ptr(),
                                             p["scaling"].data_ptr(), float(max_grad), float(self.percent_dense * extent),
                                             cls.data_ptr(), split_rank.data_ptr(), scratch.cb, None, C.byref(n_split),
                                             L.stream_ptr(dev)))
            ns = int(n_split.value)
            if samples is not None:
                samples = L.require_cuda(samples, "samples")
                if samples.shape != (num_replica * ns, 3):
                    raise RuntimeError(f"samples must be [{num_replica * ns}, 3]")
            elif ns:
                # exactly the reference's draw (:472-474): the same call on the same device consumes the same stream
                selected = (cls & 2).bool()
                stds = torch.exp(p["scaling"][selected]).repeat(num_replica, 1)
                samples = torch.normal(mean=torch.zeros((stds.size(0), 3), device=dev), std=stds, generator=generator).contiguous()
            n_out = C.c_int64(0)
            src_idx = torch.empty(4 * P, dtype=torch.int32, device=dev)
            kind = torch.empty(4 * P, dtype=torch.uint8, device=dev)
            big_ws = float(0.1 * extent) if max_screen_size is not None else -1.0
            has_z = bounding_box is not None
            zmin = float(bounding_box[2]) if has_z else 0.0
            L.check(lib.dgr_densify_plan(P, cls.data_ptr(), split_rank.data_ptr(), ns, p["xyz"].data_ptr(),
                                         p["scaling"].data_ptr(), p["quaternion"].data_ptr(), p["opacity"].data_ptr(),
                                         L.ptr(samples), float(min_opacity), big_ws, int(has_z), zmin, int(num_replica),
                                         src_idx.data_ptr(), kind.data_ptr(), scratch.cb, None, C.byref(n_out),
                                         L.stream_ptr(dev)))
            n = int(n_out.value)
            new = []
            n_par = len(PARAMS)
            for j, t in enumerate(self._every_tensor()):
                row = t.numel() // P
                dst = torch.empty((n,) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
                if n:
                    L.check(lib.dgr_gather_rows(n, row, t.data_ptr(), dst.data_ptr(), src_idx.data_ptr(), kind.data_ptr(),
                                                int(j >= n_par), L.stream_ptr(dev)))
                new.append(dst)
            if n and ns:
                L.check(lib.dgr_densify_split_fixup(n, src_idx.data_ptr(), kind.data_ptr(), split_rank.data_ptr(), ns,
                                                    p["xyz"].data_ptr(), p["scaling"].data_ptr(), p["quaternion"].data_ptr(),
                                                    samples.data_ptr(), int(num_replica), new[0].data_ptr(),
                                                    new[4].data_ptr(), L.stream_ptr(dev)))
        self._set_every_tensor(new)
        # densification_postfix (:400-404) zeroes the statistics of every row
        self.xyz_gradient_accum = torch.zeros(n, 1, device=dev)
        self.denom = torch.zeros(n, 1, device=dev)
        self.max_radii2D = torch.zeros(n, device=dev)
        return n

    # ---- pruning -----------------------------------------------------------------------------------------------
    @torch.no_grad()
    def prune_points(self, mask: torch.Tensor):
        """prune_points (:406-416): drop the rows where mask is True, Adam moments and statistics alike."""
        keep = ~mask.to(torch.bool).reshape(-1)
        ts = self._every_tensor() + [self.xyz_gradient_accum, self.denom, self.max_radii2D]
        out = compact(ts, keep.contiguous())
        self._set_every_tensor(out[:-3])
        self.xyz_gradient_accum, self.denom, self.max_radii2D = out[-3:]

    @torch.no_grad()
    def prune_gaussians(self, percent: float, import_score: torch.Tensor):
        """prune_gaussians_with_opt / prune_gaussians (:418-431): drop everything at or below the percent-quantile
        of the (LightGaussian) importance score."""
        sorted_tensor, _ = torch.sort(import_score, dim=0)
        index_nth_percentile = int(percent * (sorted_tensor.shape[0] - 1))
        value_nth_percentile = sorted_tensor[index_nth_percentile]
        self.prune_points((import_score <= value_nth_percentile).reshape(-1))

    # ---- opacity reset -----------------------------------------------------------------------------------------
    @torch.no_grad()
    def reset_opacity(self, cap: float = 0.01):
        """reset_opacity (:362-367) + replace_tensor_to_optimizer (:34-48), in place, one launch."""
        P, dev = self.num_gaussians, self.device
        m = self.exp_avg["opacity"] if self.exp_avg is not None else None
        v = self.exp_avg_sq["opacity"] if self.exp_avg_sq is not None else None
        with torch.cuda.device(dev):
            L.check(L.lib().dgr_reset_opacity(P, self.p["opacity"].data_ptr(), L.ptr(m), L.ptr(v), float(cap),
                                              L.stream_ptr(dev)))
