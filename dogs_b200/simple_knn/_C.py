"""`simple_knn._C.distCUDA2` (submodules/simple-knn/spatial.cu:17-35)."""
import torch

from .. import _lib as L


def distCUDA2(points):
    """Mean squared distance from each point to its 3 nearest neighbours; points [P,3] float32 CUDA."""
    points = L.require_cuda(points, "points")
    if points.dim() != 2 or points.size(1) != 3:
        raise RuntimeError("points must have dimensions (num_points, 3)")
    P = points.size(0)
    means = torch.zeros((P,), dtype=torch.float32, device=points.device)
    if P:
        scratch = L.Allocator(points.device)
        with torch.cuda.device(points.device):
            L.check(L.lib().sknn_dist2(P, points.data_ptr(), means.data_ptr(), scratch.cb, None,
                                       L.stream_ptr(points.device)))
    return means
