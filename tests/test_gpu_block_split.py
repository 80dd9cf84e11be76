"""Re-split of the fused model into overlapping blocks on the GPU (last bullet) against
the numpy restatement of master_gaussian_trainer.select_gaussians_in_each_block / points_in_bbox2D."""
import numpy as np
import pytest
import torch

from oracle import oracle as O
from dogs_b200 import admm as A

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _scene(n, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    xyz = torch.stack([torch.rand(n, generator=g) * 8 - 4, torch.rand(n, generator=g) * 4 - 2,
                       torch.randn(n, generator=g) * 0.1], dim=1)
    # 2 x 4 grid of unit cells expanded x1.4 about their centres (urban3d_admm.yaml:18-20); some points fall outside
    boxes = []
    for iy in range(2):
        for ix in range(4):
            cx, cy = -3.0 + 2.0 * ix * 0.9, -0.9 + 1.8 * iy * 0.9
            boxes.append([[cx - 0.7 * 1.4, cy - 0.65 * 1.4, -1.0], [cx + 0.7 * 1.4, cy + 0.65 * 1.4, 1.0]])
    th = 0.3
    M = np.array([[np.cos(th), -np.sin(th), 0.15], [np.sin(th), np.cos(th), -0.05], [0, 0, 1]], np.float64)
    return xyz, np.asarray(boxes, np.float64), M


@pytest.mark.parametrize("n,use_obb", [(1, False), (5000, False), (200_003, True)])
def test_select_gaussians_in_each_block_vs_oracle(n, use_obb):
    xyz, boxes, M = _scene(n, 4)
    M = M if use_obb else None
    valid, vis, gidx = A.select_gaussians_in_each_block(xyz.to(DEV), boxes, M)
    try:
        o_valid, o_vis, o_gidx = O.select_gaussians_in_each_block(xyz.numpy(), boxes, M)
    except ValueError:          # numpy cannot concatenate/bincount an empty selection; nothing to compare
        assert valid.numel() == 0
        return
    assert np.array_equal(valid.cpu().numpy(), o_valid)
    assert np.array_equal(vis.cpu().numpy(), o_vis)
    assert len(gidx) == len(o_gidx)
    for a, b in zip(gidx, o_gidx):
        assert np.array_equal(a.cpu().numpy(), b)
    if n > 1000:
        assert int((vis > 1).sum()) > 0 and valid.numel() < n      # overlaps and outliers both present
    # the indices feed the consensus topology directly
    if all(g.numel() for g in gidx):
        topo = A.build_topology({b: g for b, g in enumerate(gidx)}, int(valid.numel()))
        for b, g in enumerate(gidx):
            plan = topo.plans[b]
            assert plan.n_local == g.numel() and plan.S == int((vis[g] > 1).sum())
            assert torch.equal(plan.count.long(), vis[g][vis[g] > 1])


def test_fuse_block_gaussians_drops_points_outside_own_box():
    xyz, boxes, M = _scene(3000, 9)
    g = torch.Generator(device="cpu").manual_seed(1)
    blocks = []
    for b in range(3):
        idx = torch.randperm(3000, generator=g)[:800]
        x = xyz[idx].to(DEV)
        blocks.append([x] + [torch.randn(800, s, generator=g).to(DEV) for s in (3, 45, 3, 4, 1)])
    fused = A.fuse_block_gaussians(blocks, boxes[:3], M)
    exp = []
    for b in range(3):
        keep = O.points_in_bbox2d(blocks[b][0].cpu().numpy()[:, :2], boxes[b].reshape(2, 3), M)
        exp.append([t.cpu().numpy()[keep] for t in blocks[b]])
    for gi in range(6):
        assert np.array_equal(fused[gi].cpu().numpy(), np.concatenate([e[gi] for e in exp], axis=0))
