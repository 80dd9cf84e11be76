"""Prune / densify bookkeeping on the Gaussian SoA: one scan of the keep mask, one gather per tensor.

Replaces the boolean-mask indexing of `GaussianSplatModel.prune_points` and friends
(conerf/model/gaussian_fields/gaussian_splat_model.py:34-108,362-531), which runs `t[mask]` —
nonzero + gather, with a host synchronisation each — on 6 parameter tensors and 12 Adam-state
tensors.  `compact(tensors, keep)` returns tensors equal to `[t[keep] for t in tensors]`."""
import ctypes as C
from typing import List, Sequence

import torch

from . import _lib as L


@torch.no_grad()
def compact(tensors: Sequence[torch.Tensor], keep: torch.Tensor) -> List[torch.Tensor]:
    if keep.dtype != torch.bool or keep.dim() != 1:
        raise RuntimeError("keep must be a 1-D bool tensor")
    if not keep.is_cuda:
        raise RuntimeError("keep must be a CUDA tensor (dogs_b200 has no CPU path)")
    P = keep.numel()
    dev = keep.device
    keep = keep.contiguous()
    for t in tensors:
        if not t.is_cuda or t.shape[0] != P or t.element_size() != 4:
            raise RuntimeError("compact: every tensor must be a CUDA tensor with P rows and a 4-byte dtype")
    offsets = torch.empty(P, dtype=torch.int32, device=dev)
    count = C.c_int64(0)
    scratch = L.Allocator(dev)
    with torch.cuda.device(dev):
        L.check(L.lib().dgr_compact_plan(P, keep.data_ptr(), offsets.data_ptr(), scratch.cb, None,
                                         C.byref(count), L.stream_ptr(dev)))
        n = int(count.value)
        outs = []
        for t in tensors:
            src = t.contiguous()
            row = src.numel() // P if P else 0
            dst = torch.empty((n,) + tuple(src.shape[1:]), dtype=src.dtype, device=dev)
            if n and row:
                L.check(L.lib().dgr_compact_rows(P, row, src.data_ptr(), dst.data_ptr(), keep.data_ptr(),
                                                 offsets.data_ptr(), L.stream_ptr(dev)))
            outs.append(dst)
    return outs
