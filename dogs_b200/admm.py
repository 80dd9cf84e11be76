"""ADMM Gaussian consensus, one scene block per GPU (one process per GPU, torch.distributed/NCCL).

Replaces the master/slave RPC round of the reference
(conerf/trainers/master_gaussian_trainer.py:665-728, slave_gaussian_trainer.py:81-202):

  reference (star RPC through rank 0)                     here (no master)
  -----------------------------------------------------   -------------------------------------------
  send_local_model: every block ships all 59 floats/G      pack: only Gaussians owned by >= 2 blocks are
  to GPU0; z = scatter-add / count (master:538-555)        gathered into a dense [S_total,59] buffer
  broadcast_global_gaussian_splat (master:523-535)         ONE all-reduce(sum) over NVLink on that buffer
                                                           (default: neighbour exchange - each rank sends its shared
                                                           rows only to the other owners, traffic independent of N)
  update_dual_variables (slave:99-121)                     unpack_update kernel: z = sum/count, u += 1.5(x-z),
  compute_primal/dual_residual (master:396-456)            partial residual sums in the same pass
  adapt_penalty_parameters + set_penalty (master:337-393)  12-double all-reduce, identical rho update on every rank

For a Gaussian owned by a single block the reference's average is z == x and u stays 0, so it needs
no communication; the results are identical.

The device kernels live in csrc/admm.cu (C-ABI: admm_pack / admm_unpack_update / admm_penalty).
`ops` can be replaced by the tests to drive the host logic on CPU with the oracle; the default ops
call the CUDA library and raise if it is missing.
"""
import ctypes as C
from typing import List, Optional, Sequence

import torch
import torch.distributed as dist

GROUP_SIZES = (3, 3, 45, 3, 4, 1)       # xyz, f_dc, f_rest, scaling(raw), quaternion(raw), opacity(raw)
ROW = sum(GROUP_SIZES)                  # 59 floats = 236 B per Gaussian
DEFAULT_ALPHAS = (1e5, 1e4, 1e5, 1e4, 1e5, 1e4)  # config/gaussian_splatting/urban3d_admm.yaml:45-50


class ConsensusPlan:
    """Static metadata of one block (rank): which local Gaussians are shared and where they live in
    the dense exchange buffer.  Built once after the blocks are (re-)split
    (select_gaussians_in_each_block, master:124-172)."""

    def __init__(self, global_indices: torch.Tensor, count: torch.Tensor):
        n_global = count.numel()
        shared_global = count > 1
        slot_of_global = torch.cumsum(shared_global.to(torch.int64), 0) - 1
        self.n_global = int(n_global)
        self.n_local = int(global_indices.numel())
        self.S_total = int(shared_global.sum().item())
        cnt_local = count[global_indices].to(torch.float32)
        is_shared = shared_global[global_indices]
        self.slot_of_local = torch.where(is_shared, slot_of_global[global_indices],
                                         torch.full_like(global_indices, -1)).contiguous()
        self.inv_count = (1.0 / cnt_local).contiguous()
        self.local_idx = torch.nonzero(is_shared).squeeze(1).contiguous()
        self.slot = self.slot_of_local[self.local_idx].contiguous()
        self.S = int(self.local_idx.numel())
        # neighbour exchange: position of each local Gaussian in this rank's OWN compact buffer [S, 59]
        # and, per peer rank that co-owns Gaussians, the rows of that buffer exchanged with it (ordered by
        # global index, so both sides agree); filled by build_plan(exchange="neighbour")
        self.compact_slot_of_local = torch.where(
            is_shared, torch.cumsum(is_shared.to(torch.int64), 0) - 1, torch.full_like(global_indices, -1)).contiguous()
        self.peers = None

    def to(self, device):
        for k in ("slot_of_local", "inv_count", "local_idx", "slot", "compact_slot_of_local"):
            setattr(self, k, getattr(self, k).to(device))
        if self.peers is not None:
            self.peers = [(r, rows.to(device)) for r, rows in self.peers]
        return self

    def compact(self):
        """The same plan addressed through the rank-local buffer (slot = 0..S-1) instead of the global
        dense one: what pack / unpack_update see in neighbour-exchange mode."""
        import copy
        p = copy.copy(self)
        p.slot_of_local = self.compact_slot_of_local
        p.slot = torch.arange(self.S, dtype=torch.int64, device=self.local_idx.device)
        p.S_total = self.S
        return p


def build_plan(global_indices: torch.Tensor, n_global: int, group=None, exchange: str = "neighbour") -> ConsensusPlan:
    """count[i] = number of blocks holding global Gaussian i (torch.bincount at master:168),
    obtained with one all-reduce instead of on a master.
    exchange="neighbour" (default) also finds, once, which rows are co-owned with which rank, so that a
    consensus round sends every shared row only to its other owners (point-to-point, S_k * 236 B per rank
    whatever the number of blocks); exchange="dense" keeps the single all-reduce over the [S_total, 59]
    buffer of ALL shared Gaussians, which grows with the number of blocks (1.2 GB at 8 x 3 M)."""
    count = torch.bincount(global_indices, minlength=n_global).to(torch.int32)
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    if multi:
        dist.all_reduce(count, op=dist.ReduceOp.SUM, group=group)
    plan = ConsensusPlan(global_indices, count)
    if multi and exchange == "neighbour":
        plan.peers = _find_peers(global_indices, n_global, plan, group)
    return plan


def _find_peers(global_indices, n_global, plan, group):
    """[(peer rank in group, rows of the local compact buffer co-owned with it, ordered by global index)]"""
    dev = global_indices.device
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    n = torch.tensor([global_indices.numel()], dtype=torch.int64, device=dev)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    pad = torch.full((max(max(sizes), 1),), -1, dtype=torch.int64, device=dev)
    pad[:global_indices.numel()] = global_indices
    gathered = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(gathered, pad, group=group)
    peers = []
    mark = torch.zeros(n_global, dtype=torch.bool, device=dev)
    for r in range(world):
        if r == rank or sizes[r] == 0:
            continue
        theirs = gathered[r][:sizes[r]]
        mark.zero_()
        mark[theirs] = True
        rows_local = torch.nonzero(mark[global_indices]).squeeze(1)
        if rows_local.numel() == 0:
            continue
        rows_local = rows_local[torch.argsort(global_indices[rows_local])]
        peers.append((r, plan.compact_slot_of_local[rows_local].contiguous()))
    return peers


class CudaOps:
    """The product path: sm_100a kernels behind the C-ABI."""

    def __init__(self):
        from . import _lib
        self.L = _lib
        self.lib = _lib.lib()      # raises if the library has not been built

    def _t6(self, ts):
        arr = self.L.AdmmTensors()
        for i, t in enumerate(ts):
            if t is None:            # group handled elsewhere (fused into the rasterizer backward)
                arr.g[i] = None
                continue
            if not (t.is_cuda and t.is_contiguous() and t.dtype == torch.float32):
                raise RuntimeError("ADMM tensors must be contiguous float32 CUDA tensors")
            arr.g[i] = t.data_ptr()
        return arr

    def pack(self, x, plan, sendbuf):
        xs = self._t6(x)
        dev = sendbuf.device
        with torch.cuda.device(dev):
            self.L.check(self.lib.admm_pack(C.byref(xs), plan.local_idx.data_ptr(), plan.slot.data_ptr(),
                                            plan.S, plan.S_total, 1, sendbuf.data_ptr(),
                                            self.L.stream_ptr(dev)))

    def unpack_update(self, x, u, z, plan, recvbuf, relax, residuals, zmu=None):
        xs, us, zs = self._t6(x), self._t6(u), self._t6(z)
        zm = self._t6(zmu) if zmu is not None else None
        dev = residuals.device
        with torch.cuda.device(dev):
            self.L.check(self.lib.admm_unpack_update(
                C.byref(xs), C.byref(us), C.byref(zs), C.byref(zm) if zm is not None else None,
                plan.slot_of_local.data_ptr(),
                plan.inv_count.data_ptr(), plan.n_local, recvbuf.data_ptr() if recvbuf.numel() else None,
                float(relax), residuals.data_ptr(), self.L.stream_ptr(dev)))

    def penalty(self, x, u, z, grads, n, coef, sums):
        xs, zs = self._t6(x), self._t6(z)
        us = self._t6(u) if u is not None else None
        gs = self._t6(grads) if grads is not None else None
        carr = (C.c_float * 6)(*[float(c) for c in coef])
        dev = sums.device
        with torch.cuda.device(dev):
            self.L.check(self.lib.admm_penalty(C.byref(xs), C.byref(us) if us is not None else None, C.byref(zs),
                                               C.byref(gs) if gs is not None else None, n, carr,
                                               sums.data_ptr(), self.L.stream_ptr(dev)))


class AdmmBlock:
    """ADMM state of one block: x (the model's raw parameter tensors, shared with the optimizer),
    scaled duals u, local copy z of the consensus variables, penalties rho."""

    def __init__(self, params: Sequence[torch.Tensor], plan: ConsensusPlan,
                 alphas: Sequence[float] = DEFAULT_ALPHAS, over_relaxation_coeff: float = 0.5,
                 mu: float = 10.0, tau_inc: float = 2.0, tau_dec: float = 2.0, group=None,
                 ops=None):
        assert len(params) == 6
        self.x: List[torch.Tensor] = list(params)
        n = plan.n_local
        for t, sz in zip(self.x, GROUP_SIZES):
            assert t.numel() == n * sz, "parameter tensor does not match the plan"
        self.plan = plan
        self.group = group
        self.ops = ops if ops is not None else CudaOps()
        dev = self.x[0].device
        # setup_dual_variables (slave:81-97): u := 0;  z := x (set_global_gaussians after fusion)
        self.u = [torch.zeros_like(t) for t in self.x]
        self.z = [t.detach().clone() for t in self.x]
        # zmu = z - u, refreshed by every consensus round: the per-iteration penalty then reads one
        # tensor set instead of two (r = x - zmu)
        self.zmu = [t.detach().clone() for t in self.x]
        # setup_penalty_parameters (master:326-334): rho_g = alpha_g / N_global
        self.rho = [a / plan.n_global for a in alphas]
        self.relax = 1.0 + over_relaxation_coeff
        self.mu, self.tau_inc, self.tau_dec = mu, tau_inc, tau_dec
        # exchange buffer: the rank's own shared rows (neighbour exchange) or all shared rows (dense all-reduce)
        self.neighbour = plan.peers is not None
        self.xplan = plan.compact() if self.neighbour else plan
        self.buf = torch.zeros(max(self.xplan.S_total, 1) * ROW, dtype=torch.float32, device=dev)
        self._send = [torch.empty(rows.numel(), ROW, dtype=torch.float32, device=dev) for _, rows in (plan.peers or [])]
        self._recv = [torch.empty_like(s) for s in self._send]
        self.residuals = torch.zeros(12, dtype=torch.float64, device=dev)
        self.pen_sums = torch.zeros(6, dtype=torch.float64, device=dev)
        self.fused_sums = torch.zeros(6, dtype=torch.float64, device=dev)
        self.fused = False
        self.fused_all = False
        self._pen_scale, self._pen_scale_key = None, None
        self.last_primal: Optional[List[float]] = None
        self.last_dual: Optional[List[float]] = None

    def _world(self):
        if dist.is_available() and dist.is_initialized():
            return dist.get_world_size(self.group)
        return 1

    @torch.no_grad()
    def consensus(self, adapt: bool = True):
        """One consensus round: steps (2)-(6) of master_gaussian_trainer.train_iteration."""
        x = [t.detach() for t in self.x]
        self.ops.pack(x, self.xplan, self.buf)
        if self.neighbour:
            self._exchange_with_peers()
        elif self._world() > 1 and self.plan.S_total > 0:
            dist.all_reduce(self.buf, op=dist.ReduceOp.SUM, group=self.group)
        self.residuals.zero_()
        self.ops.unpack_update(x, self.u, self.z, self.xplan, self.buf, self.relax, self.residuals, self.zmu)
        res = self.residuals.clone()
        # primal_g = sum_k mse(z[idx_k], x_k): every block divides by its own element count
        n = self.plan.n_local
        for g, sz in enumerate(GROUP_SIZES):
            res[g] /= max(n * sz, 1)
        if self._world() > 1:
            dist.all_reduce(res, op=dist.ReduceOp.SUM, group=self.group)
        res = res.tolist()
        primal = res[:6]
        dual = [self.rho[g] * res[6 + g] / (self.plan.n_global * GROUP_SIZES[g]) for g in range(6)]
        self.last_primal, self.last_dual = primal, dual
        if adapt:  # adapt_penalty_parameters (master:337-377), evaluated identically on every rank
            for g in range(6):
                if primal[g] > self.mu * dual[g]:
                    self.rho[g] *= self.tau_inc
                elif dual[g] > self.mu * primal[g]:
                    self.rho[g] /= self.tau_dec
        return primal, dual

    def _exchange_with_peers(self):
        """buf[row] += the same Gaussian's parameters from every other owner: one send + one receive per
        co-owning rank (batched point-to-point), instead of an all-reduce over all shared Gaussians."""
        rows = self.buf.view(-1, ROW)
        ops = []
        for (r, idx), s, rv in zip(self.plan.peers, self._send, self._recv):
            torch.index_select(rows, 0, idx, out=s)                  # before anything is added
            peer = dist.get_global_rank(self.group, r) if self.group is not None else r
            ops.append(dist.P2POp(dist.isend, s, peer, self.group))
            ops.append(dist.P2POp(dist.irecv, rv, peer, self.group))
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        for (_, idx), rv in zip(self.plan.peers, self._recv):       # fixed peer order => deterministic sums
            rows.index_add_(0, idx, rv)

    def fused_coef(self):
        n = self.plan.n_local
        return [self.rho[g] / max(n * GROUP_SIZES[g], 1) for g in range(6)]

    def fuse_into_rasterizer(self, enable: bool = True, all_groups: bool = False):
        """Let the rasterizer backward add the penalty gradient of xyz / f_dc / f_rest (51 of the 59
        floats) while it writes those gradients anyway (DgrAdmmPenalty); add_penalty_grads_ then only
        touches scaling / quaternion / opacity.  Valid when the rasterizer is fed this block's raw
        xyz, f_dc and f_rest tensors (conerf's render() does).
        all_groups=True also fuses scaling / quaternion / opacity: valid when the rasterizer receives
        the consensus (raw) tensors themselves, i.e. GaussianRasterizationSettings.raw_activations;
        add_penalty_grads_ then launches no penalty kernel at all."""
        from . import diff_gaussian_rasterization as dgr
        self.fused = bool(enable)
        self.fused_all = bool(enable and all_groups)
        dgr.set_fused_admm(self if enable else None)
        self.fused_sums.zero_()

    @torch.no_grad()
    def add_penalty_grads_(self):
        """add_admm_penalties (slave:161-202) fused with its own backward: accumulates
        rho_g/numel_g * (x+u-z) into x.grad and returns the 6 penalty values as a device tensor
        (no .item() host syncs, unlike slave:195-200)."""
        n = self.plan.n_local
        coef = [self.rho[g] / max(n * GROUP_SIZES[g], 1) for g in range(6)]
        for t in self.x:
            if t.grad is None:
                t.grad = torch.zeros_like(t)
        xs = [t.detach() for t in self.x]
        if self.fused_all:   # every group was handled by the rasterizer backward of this iteration
            self.pen_sums.copy_(self.fused_sums)
            self.fused_sums.zero_()
        else:
            self.pen_sums.zero_()
            if self.fused:   # groups 0-2 were handled by the rasterizer backward of this iteration
                xs = [None, None, None] + xs[3:]
            self.ops.penalty(xs, None, self.zmu, [t.grad for t in self.x], n, coef, self.pen_sums)
            if self.fused:
                self.pen_sums[:3] += self.fused_sums[:3]
                self.fused_sums.zero_()
        # penalties_g = 0.5 * rho_g/numel_g * sum r^2; the scale vector lives on the device and is only
        # rebuilt when rho changes (no host->device copy, no sync, in the per-iteration path)
        key = tuple(coef)
        if self._pen_scale_key != key:
            self._pen_scale = torch.tensor([0.5 * c for c in coef], dtype=torch.float64,
                                           device=self.pen_sums.device)
            self._pen_scale_key = key
        return self.pen_sums * self._pen_scale

    def state_dict(self):
        """What a checkpoint must persist for ADMM (slave:209-263 stores u_*, rho_*; the reference does
        not persist z / indices)."""
        return dict(u=self.u, z=self.z, rho=list(self.rho), slot_of_local=self.plan.slot_of_local,
                    inv_count=self.plan.inv_count)


# ---- fusion / re-split of the blocks (once per run, before ADMM starts) ----------------------------
def _boxes2d(bboxes, device):
    """[n_blocks, 2, 3] (or [n_blocks, 6]) min/max corners as the reference stores them -> [n_blocks, 4]
    float64 {min_x, min_y, max_x, max_y} (points_in_bbox2D only looks at x and y)."""
    b = torch.as_tensor(bboxes, dtype=torch.float64).reshape(-1, 2, 3)
    return torch.cat([b[:, 0, :2], b[:, 1, :2]], dim=1).contiguous().to(device)


def block_membership(xyz: torch.Tensor, bboxes, world_to_obb=None):
    """(mask int64 [N], count int32 [N]): bit b of mask[i] = Gaussian i lies in block b's box
    (points_in_bbox2D, conerf/datasets/utils.py:186-205, for all blocks in one pass on the GPU)."""
    from . import _lib as L
    xyz = L.require_cuda(xyz.detach(), "xyz")
    n = xyz.shape[0]
    boxes = _boxes2d(bboxes, xyz.device)
    mask = torch.zeros(n, dtype=torch.int64, device=xyz.device)
    count = torch.zeros(n, dtype=torch.int32, device=xyz.device)
    t = None
    if world_to_obb is not None:
        t = (C.c_double * 9)(*[float(v) for v in torch.as_tensor(world_to_obb, dtype=torch.float64).reshape(-1)])
    with torch.cuda.device(xyz.device):
        L.check(L.lib().admm_block_membership(n, xyz.data_ptr(), t, boxes.shape[0], boxes.data_ptr(),
                                              mask.data_ptr(), count.data_ptr(), L.stream_ptr(xyz.device)))
    return mask, count


def select_gaussians_in_each_block(xyz: torch.Tensor, bboxes, world_to_obb=None):
    """GPU restatement of master_gaussian_trainer.select_gaussians_in_each_block (:124-172).
    Returns (valid_gaussian_indices, visibility_count, global_indices):
      valid_gaussian_indices  int64 [N'] Gaussians inside at least one box (feed to extract_sub_gaussians),
      visibility_count        int64 [N'] number of blocks holding each kept Gaussian (the bincount),
      global_indices          list over blocks of int64 index tensors into the kept Gaussians, ascending."""
    mask, count = block_membership(xyz, bboxes, world_to_obb)
    valid = torch.nonzero(count).squeeze(-1)
    mask, count = mask[valid], count[valid]
    n_blocks = _boxes2d(bboxes, "cpu").shape[0]
    global_indices = [torch.nonzero((mask >> b) & 1).squeeze(-1) for b in range(n_blocks)]
    return valid, count.to(torch.int64), global_indices


def fuse_block_gaussians(block_params: Sequence[Sequence[torch.Tensor]], point_bboxes=None, world_to_obb=None):
    """GPU restatement of master_gaussian_trainer.fuse_block_gaussians (:37-105) without its .ply dumps:
    every block first drops the Gaussians outside its own initial box, then the six parameter tensors
    (xyz, f_dc, f_rest, raw scaling, raw quaternion, raw opacity) are concatenated block by block."""
    kept = []
    for b, params in enumerate(block_params):
        params = list(params)
        if point_bboxes is not None:
            box = torch.as_tensor(point_bboxes[b], dtype=torch.float64).reshape(1, 2, 3)
            mask, _ = block_membership(params[0], box, world_to_obb)
            idx = torch.nonzero(mask).squeeze(-1)
            params = [t.detach()[idx] for t in params]
        kept.append(params)
    return [torch.cat([k[g] for k in kept], dim=0) for g in range(6)]
