"""ADMM Gaussian consensus over scene blocks (one process per GPU, any number of blocks per process).

Replaces the master/slave RPC round of the reference
(conerf/trainers/master_gaussian_trainer.py:665-728, slave_gaussian_trainer.py:81-202):

  reference (star RPC through rank 0)                     here (no master)
  -----------------------------------------------------   -------------------------------------------
  send_local_model: every block ships all 59 floats/G      pack_rows: only Gaussians owned by >= 2 blocks are
  to GPU0; z = scatter-add / count (master:538-555)        gathered, as 236-byte rows, straight into the
  broadcast_global_gaussian_splat (master:523-535)         buffers exchanged with their co-owners: NCCL
                                                           send/recv between GPUs (S_k * 236 B per block whatever
                                                           the number of blocks), zero-copy between blocks of one GPU
  update_dual_variables (slave:99-121)                     consensus_update kernel: z = sum/count with the owners
  compute_primal/dual_residual (master:396-456)            added in the reference's block order (bit-identical z),
                                                           u += 1.5(x-z), residual sums, in one pass
  adapt_penalty_parameters + set_penalty (master:337-393)  12-double all-reduce, identical rho update on every rank

For a Gaussian owned by a single block the reference's average is z == x and u stays 0, so it needs
no communication; the results are identical.

The device kernels live in csrc/admm.cu (C-ABI: admm_pack_rows / admm_consensus_update / admm_penalty).
`ops` can be replaced by the tests to drive the host logic on CPU; the default ops call the CUDA
library and raise if it is missing.
"""
import ctypes as C
import os
from typing import Dict, List, Optional, Sequence

import torch
import torch.distributed as dist

GROUP_SIZES = (3, 3, 45, 3, 4, 1)       # xyz, f_dc, f_rest, scaling(raw), quaternion(raw), opacity(raw)
GROUP_KEYS = ("xyz", "fdc", "fr", "s", "q", "o")      # suffixes of the reference's checkpoint keys (slave:209-263)
ROW = sum(GROUP_SIZES)                  # 59 floats = 236 B per Gaussian
DEFAULT_ALPHAS = (1e5, 1e4, 1e5, 1e4, 1e5, 1e4)  # config/gaussian_splatting/urban3d_admm.yaml:45-50
SELF_ADDEND = 15 << 28                  # addend-table entry "this block's own x" (include/dogs_b200.h)
NO_ADDEND = 0xFFFFFFFF
SRC_SEND, SRC_RECV = 0, 1               # exchange buffers a consensus kernel reads addends from


def _dist_on(group=None):
    return dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1


class BlockPlan:
    """Static consensus metadata of one local block (built once after the blocks are (re-)split,
    select_gaussians_in_each_block, master:124-172)."""

    def __init__(self, block_id, n_local, n_global):
        self.block_id, self.n_local, self.n_global = int(block_id), int(n_local), int(n_global)
        self.shared_row = None      # int32 [n_local]: row of the addend table, -1 = owned by this block only
        self.addends = None         # int64 on host side of the build, stored as int32 bit patterns [S, max_addends]
        self.count = None           # uint8 [S] number of owners
        self.max_addends = 0
        self.S = 0
        self.pack_src = None        # int32 [R_a] local Gaussian of every row this block sends
        self.pack_dst = None        # dense exchange only: int32 [S] row of the all-reduced buffer
        self.send_start = 0         # first row of this block's region in the rank's send buffer
        self.send_rows = 0
        self.sends = []             # [(dest rank, first row, rows)] contiguous regions of the send buffer
        self.peers = []             # co-owning block ids (diagnostics)
        self.messages = []          # [(dest block, dest rank, first row in the send buffer, rows)] in pack order
        self.p2p_sel = None         # peer-memory exchange: uint8 [R_a] destination buffer of every packed row
        self.p2p_row = None         #                       int32 [R_a] row in that buffer

    def to(self, device):
        for k in ("shared_row", "addends", "count", "pack_src", "pack_dst", "p2p_sel", "p2p_row"):
            if getattr(self, k) is not None:
                setattr(self, k, getattr(self, k).to(device))
        return self


class Topology:
    """Which block lives where, and the layout of the exchange buffers of THIS rank."""

    def __init__(self):
        self.rank, self.world = 0, 1
        self.n_global = 0
        self.owner_rank: Dict[int, int] = {}
        self.plans: Dict[int, BlockPlan] = {}
        self.send_rows_total = 0
        self.recv_rows_total = 0
        self.recvs = []             # [(src rank, first row, rows)] regions of the recv buffer, in posting order
        self.exchange = "neighbour"
        self.S_total = 0            # dense mode: rows of the all-reduced buffer
        self.group = None
        self.msg_recv = {}          # (sender block, receiver block) -> first row in THIS rank's receive buffer

    def to(self, device):
        for p in self.plans.values():
            p.to(device)
        return self

    # single-block conveniences (one block per rank, the reference's deployment)
    @property
    def only(self) -> BlockPlan:
        assert len(self.plans) == 1, "topology holds several local blocks"
        return next(iter(self.plans.values()))


def _gather_blocks(local_blocks, group):
    """{block id: (owner rank, int64 global indices on the local device)} for ALL blocks of the job."""
    if not _dist_on(group):
        return {int(b): (0, g) for b, g in local_blocks.items()}
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = next(iter(local_blocks.values())).device if local_blocks else torch.device("cpu")
    ids = sorted(int(b) for b in local_blocks)
    meta = torch.full((64, 2), -1, dtype=torch.int64, device=dev)          # (block id, n) per local block
    assert len(ids) <= 64, "at most 64 blocks per rank"
    for j, b in enumerate(ids):
        meta[j, 0], meta[j, 1] = b, local_blocks[b].numel()
    metas = [torch.empty_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta, group=group)
    metas = [m.cpu() for m in metas]
    totals = [int(m[:, 1].clamp(min=0).sum()) for m in metas]
    pad = torch.full((max(max(totals), 1),), -1, dtype=torch.int64, device=dev)
    if ids:
        cat = torch.cat([local_blocks[b].reshape(-1).to(torch.int64) for b in ids])
        pad[:cat.numel()] = cat
    gathered = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(gathered, pad, group=group)
    out = {}
    for r in range(world):
        off = 0
        for j in range(64):
            b, n = int(metas[r][j, 0]), int(metas[r][j, 1])
            if b < 0:
                break
            assert b not in out, f"block {b} is held by two ranks"
            out[b] = (r, local_blocks[b].reshape(-1).to(torch.int64) if r == rank else gathered[r][off:off + n])
            off += n
    return out


def build_topology(local_blocks: Dict[int, torch.Tensor], n_global: int, group=None,
                   exchange: str = "neighbour") -> Topology:
    """local_blocks: {block id: global indices of its Gaussians (global_indices[k], master:124-172)} for the
    blocks trained by THIS process.  Collective over `group` (every rank calls it once).

    exchange="neighbour" (default): every shared Gaussian travels only to its other owners, in 59-float rows
      ordered by global index on both sides; addends are listed in ascending block order, which makes z
      bit-identical to the reference's master-side average.
    exchange="dense": the literal "one all-reduce over the [S_total, 59] buffer of all shared Gaussians"
      (one block per rank only; grows with the number of blocks, and NCCL chooses the summation order)."""
    topo = Topology()
    topo.group, topo.exchange, topo.n_global = group, exchange, int(n_global)
    if _dist_on(group):
        topo.rank, topo.world = dist.get_rank(group), dist.get_world_size(group)
    allb = _gather_blocks(local_blocks, group)
    ids = sorted(allb)
    topo.owner_rank = {b: allb[b][0] for b in ids}
    dev = next(iter(local_blocks.values())).device if local_blocks else torch.device("cpu")
    count = torch.zeros(n_global, dtype=torch.int32, device=dev)
    for b in ids:
        g = allb[b][1]
        if g.numel() and int(torch.bincount(g, minlength=1).max()) > 1:
            raise ValueError(f"block {b} lists a Gaussian twice")
        count[g] += 1                                                     # torch.bincount of master:168
    if n_global and int(count.max()) > 255:
        raise ValueError("a Gaussian is owned by more than 255 blocks")
    local_ids = sorted(int(b) for b in local_blocks)

    if exchange == "dense":
        if len(local_ids) != 1:
            raise ValueError('exchange="dense" supports one block per rank')
        a = local_ids[0]
        ga = allb[a][1]
        shared_global = count > 1
        slot_of_global = torch.cumsum(shared_global.to(torch.int64), 0) - 1
        plan = BlockPlan(a, ga.numel(), n_global)
        is_shared = shared_global[ga]
        plan.S = int(is_shared.sum())
        plan.shared_row = torch.where(is_shared, torch.cumsum(is_shared.to(torch.int64), 0) - 1,
                                      torch.full_like(ga, -1)).to(torch.int32)
        loc = torch.nonzero(is_shared).squeeze(1)
        plan.pack_src = loc.to(torch.int32)
        plan.pack_dst = slot_of_global[ga[loc]].to(torch.int32)           # row of the dense buffer
        plan.addends = _as_i32(((SRC_SEND << 28) | slot_of_global[ga[loc]]).reshape(-1, 1))
        plan.count = count[ga[loc]].to(torch.uint8)
        plan.max_addends = 1
        plan.send_rows = plan.S
        topo.S_total = int(shared_global.sum())
        topo.send_rows_total = topo.S_total
        topo.plans[a] = plan
        return topo
    if exchange != "neighbour":
        raise ValueError(f"unknown exchange {exchange!r}")

    # ---- neighbour exchange ------------------------------------------------------------------------------
    # common[(a, b)] for a < b: (rows in a, rows in b) of the Gaussians both own, ordered by global index
    pos = {}
    for b in ids:
        p = torch.full((n_global,), -1, dtype=torch.int64, device=dev)
        g = allb[b][1]
        p[g] = torch.arange(g.numel(), dtype=torch.int64, device=dev)
        pos[b] = p
    involved = set(local_ids)
    common = {}
    for i, a in enumerate(ids):
        for b in ids[i + 1:]:
            if a not in involved and b not in involved:
                continue
            gb = allb[b][1]
            both = gb[pos[a][gb] >= 0]
            if both.numel() == 0:
                continue
            both = torch.sort(both).values
            common[(a, b)] = (pos[a][both], pos[b][both])

    def rows_of(a, b):       # local rows of block a shared with block b (same order on both sides), or None
        key = (a, b) if a < b else (b, a)
        if key not in common:
            return None
        return common[key][0 if a < b else 1]

    # send buffer: for each local block a (ascending): for each dest rank (ascending): for each block b there
    msg_send = {}            # (a, b) -> first row of message a->b in the send buffer
    row = 0
    for a in local_ids:
        plan = BlockPlan(a, allb[a][1].numel(), n_global)
        plan.send_start = row
        srcs = []
        for r in range(topo.world):
            first = row
            for b in ids:
                if b == a or topo.owner_rank[b] != r:
                    continue
                ra = rows_of(a, b)
                if ra is None:
                    continue
                msg_send[(a, b)] = row
                plan.messages.append((b, r, row, int(ra.numel())))
                srcs.append(ra)
                row += ra.numel()
                plan.peers.append(b)
            if r != topo.rank and row > first:
                plan.sends.append((r, first, row - first))
        plan.send_rows = row - plan.send_start
        plan.pack_src = (torch.cat(srcs) if srcs else torch.zeros(0, dtype=torch.int64, device=dev)).to(torch.int32)
        plan.pack_dst = None
        topo.plans[a] = plan
    topo.send_rows_total = row
    # recv buffer: for each source rank (ascending): for each sender block b there (ascending): for each local a
    msg_recv = {}            # (b, a) -> first row of message b->a in the recv buffer
    row = 0
    for r in range(topo.world):
        if r == topo.rank:
            continue
        for b in ids:
            if topo.owner_rank[b] != r:
                continue
            first = row
            for a in local_ids:
                ra = rows_of(a, b)
                if ra is None:
                    continue
                msg_recv[(b, a)] = row
                row += ra.numel()
            if row > first:
                topo.recvs.append((r, first, row - first))
    topo.recv_rows_total = row
    topo.msg_recv = dict(msg_recv)
    if max(topo.send_rows_total, topo.recv_rows_total) >= (1 << 28):
        raise ValueError("exchange buffer exceeds 2^28 rows")

    # addend tables: owners in ascending block order (plus_gaussians loop of master:549-552)
    for a in local_ids:
        plan = topo.plans[a]
        ga = allb[a][1]
        cnt_local = count[ga]
        is_shared = cnt_local > 1
        plan.S = int(is_shared.sum())
        shared_row = torch.where(is_shared, torch.cumsum(is_shared.to(torch.int64), 0) - 1, torch.full_like(ga, -1))
        plan.shared_row = shared_row.to(torch.int32)
        plan.count = cnt_local[is_shared].to(torch.uint8)
        plan.max_addends = int(cnt_local.max()) if (plan.S and ga.numel()) else 0
        tab = torch.full((plan.S, max(plan.max_addends, 1)), NO_ADDEND, dtype=torch.int64, device=dev)
        fill = torch.zeros(plan.S, dtype=torch.int64, device=dev)
        for b in ids:
            if b == a:
                rows = torch.nonzero(is_shared).squeeze(1)
                ent = torch.full((rows.numel(),), SELF_ADDEND, dtype=torch.int64, device=dev)
            else:
                rows = rows_of(a, b)
                if rows is None:
                    continue
                j = torch.arange(rows.numel(), dtype=torch.int64, device=dev)
                if topo.owner_rank[b] == topo.rank:
                    ent = (SRC_SEND << 28) | (msg_send[(b, a)] + j)     # read the co-located sender's buffer
                else:
                    ent = (SRC_RECV << 28) | (msg_recv[(b, a)] + j)
            sr = shared_row[rows]
            tab[sr, fill[sr]] = ent
            fill[sr] += 1
        assert plan.S == 0 or bool((fill == cnt_local[is_shared].to(torch.int64)).all())
        plan.addends = _as_i32(tab)
    return topo


def _as_i32(t):
    """uint32 bit patterns held in an int64 tensor -> int32 tensor with the same low 32 bits."""
    t = t & 0xFFFFFFFF
    return torch.where(t >= (1 << 31), t - (1 << 32), t).to(torch.int32).contiguous()


def build_plan(global_indices: torch.Tensor, n_global: int, group=None, exchange: str = "neighbour") -> Topology:
    """One block per rank (block id = rank), the reference's deployment."""
    rank = dist.get_rank(group) if _dist_on(group) else 0
    return build_topology({rank: global_indices}, n_global, group=group, exchange=exchange)


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

    def new_scratch(self, device):
        return torch.zeros(int(self.lib.admm_consensus_scratch_bytes()), dtype=torch.uint8, device=device)

    def pack_rows(self, x, plan, out_rows):
        """out_rows: [R, 59] view of this block's region of the send buffer."""
        if plan.send_rows == 0:
            return
        if plan.pack_dst is not None:
            raise RuntimeError("dense exchange is packed by pack_rows_scatter")
        xs = self._t6(x)
        dev = out_rows.device
        with torch.cuda.device(dev):
            self.L.check(self.lib.admm_pack_rows(C.byref(xs), plan.pack_src.data_ptr(), plan.pack_src.numel(),
                                                 out_rows.data_ptr(), self.L.stream_ptr(dev)))

    def pack_rows_to(self, x, plan, dst_ptrs):
        """Rows straight into the buffers their consumers read (own send buffer / peers' receive buffers)."""
        if plan.send_rows == 0:
            return
        xs = self._t6(x)
        dev = plan.pack_src.device
        arr = (C.c_void_p * len(dst_ptrs))(*dst_ptrs)
        with torch.cuda.device(dev):
            self.L.check(self.lib.admm_pack_rows_to(C.byref(xs), plan.pack_src.data_ptr(), plan.p2p_sel.data_ptr(),
                                                    plan.p2p_row.data_ptr(), plan.pack_src.numel(), arr, len(dst_ptrs),
                                                    self.L.stream_ptr(dev)))

    def consensus_update(self, x, u, z, zmu, plan, sources, relax, scratch, residuals):
        xs, us, zs = self._t6(x), self._t6(u), self._t6(z)
        zm = self._t6(zmu) if zmu is not None else None
        dev = residuals.device
        ptrs = (C.c_void_p * len(sources))(*[s.data_ptr() if s is not None and s.numel() else None for s in sources])
        with torch.cuda.device(dev):
            self.L.check(self.lib.admm_consensus_update(
                C.byref(xs), C.byref(us), C.byref(zs), C.byref(zm) if zm is not None else None,
                plan.shared_row.data_ptr(), plan.addends.data_ptr() if plan.S else None,
                plan.count.data_ptr() if plan.S else None, plan.max_addends, plan.n_local,
                ptrs, len(sources), float(relax), scratch.data_ptr(), residuals.data_ptr(), self.L.stream_ptr(dev)))

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
    scaled duals u, local copy z of the consensus variables (and zmu = z - u for the penalty).
    The penalties rho are global and live on the ConsensusGroup."""

    def __init__(self, params: Sequence[torch.Tensor], plan, alphas: Sequence[float] = DEFAULT_ALPHAS,
                 over_relaxation_coeff: float = 0.5, mu: float = 10.0, tau_inc: float = 2.0, tau_dec: float = 2.0,
                 group=None, ops=None, _owner=None):
        assert len(params) == 6
        topo = None
        if isinstance(plan, Topology):
            topo, plan = plan, plan.only
        self.x: List[torch.Tensor] = list(params)
        n = plan.n_local
        for t, sz in zip(self.x, GROUP_SIZES):
            assert t.numel() == n * sz, "parameter tensor does not match the plan"
        self.plan = plan
        # setup_dual_variables (slave:81-97): u := 0;  z := x (set_global_gaussians after fusion)
        self.u = [torch.zeros_like(t) for t in self.x]
        self.z = [t.detach().clone() for t in self.x]
        # zmu = z - u, refreshed by every consensus round: the per-iteration penalty then reads one
        # tensor set instead of two (r = x - zmu)
        self.zmu = [t.detach().clone() for t in self.x]
        dev = self.x[0].device
        self.residuals = torch.zeros(12, dtype=torch.float64, device=dev)
        self.pen_sums = torch.zeros(6, dtype=torch.float64, device=dev)
        self.fused_sums = torch.zeros(6, dtype=torch.float64, device=dev)
        self.fused = False
        self.fused_all = False
        self._pen_scale, self._pen_scale_key = None, None
        self.admm_enabled = True
        if _owner is None:     # stand-alone construction: one block per rank
            assert topo is not None, "pass the Topology returned by build_plan / build_topology"
            _owner = ConsensusGroup({plan.block_id: self}, topo, alphas=alphas,
                                    over_relaxation_coeff=over_relaxation_coeff, mu=mu, tau_inc=tau_inc,
                                    tau_dec=tau_dec, ops=ops)
        self.owner = _owner

    # the penalties are shared by all blocks
    @property
    def rho(self):
        return self.owner.rho

    @property
    def ops(self):
        return self.owner.ops

    @property
    def last_primal(self):
        return self.owner.last_primal

    @property
    def last_dual(self):
        return self.owner.last_dual

    def consensus(self, adapt: bool = True):
        return self.owner.consensus(adapt=adapt)

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
        add_penalty_grads_ then launches no penalty kernel at all.

        The block is handed to the rasterizer PER CALL: pass it as GaussianRasterizationSettings.admm of
        the ONE render of the iteration that shall carry the penalty (several blocks per process, several
        renders per iteration and other models of the same size are then unambiguous)."""
        self.fused = bool(enable)
        self.fused_all = bool(enable and all_groups)
        self.fused_sums.zero_()
        return self

    @torch.no_grad()
    def add_penalty_grads_(self):
        """add_admm_penalties (slave:161-202) fused with its own backward: accumulates
        rho_g/numel_g * (x+u-z) into x.grad and returns the 6 penalty values as a device tensor
        (no .item() host syncs, unlike slave:195-200)."""
        n = self.plan.n_local
        coef = self.fused_coef()
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

    # ---- checkpointing: the reference's own keys (slave_gaussian_trainer.py:209-263) ---------------------
    def state_dict(self):
        """meta_data entries the reference's slave writes (u_xyz .. u_o in the model's tensor shapes,
        rho_xyz .. rho_o, admm_enabled) plus the local consensus copy z_* (the reference re-receives it
        from the master at the next round; it is needed here to resume the penalty between rounds)."""
        sd = {"admm_enabled": bool(self.admm_enabled)}
        for k, u, z, x, rho in zip(GROUP_KEYS, self.u, self.z, self.x, self.rho):
            sd[f"u_{k}"] = u.detach().clone().view(x.shape)
            sd[f"z_{k}"] = z.detach().clone().view(x.shape)
            sd[f"rho_{k}"] = float(rho)
        return sd

    @torch.no_grad()
    def load_state_dict(self, sd):
        """Accepts a reference slave checkpoint's meta_data (u_*, rho_*, admm_enabled) or state_dict()'s
        output.  Without z_* the consensus copy restarts at z = x, as after the fusion hand-off."""
        self.admm_enabled = bool(sd.get("admm_enabled", True))
        for g, k in enumerate(GROUP_KEYS):
            u = sd[f"u_{k}"]
            if u.numel() != self.u[g].numel():
                raise ValueError(f"u_{k}: {tuple(u.shape)} does not match the block ({self.u[g].numel()} elements)")
            self.u[g].copy_(u.reshape(self.u[g].shape).to(self.u[g].device, self.u[g].dtype))
            z = sd.get(f"z_{k}")
            src = z.reshape(self.z[g].shape).to(self.z[g].device, self.z[g].dtype) if z is not None else self.x[g].detach()
            self.z[g].copy_(src)
            torch.sub(self.z[g], self.u[g], out=self.zmu[g])
            if f"rho_{k}" in sd:
                self.owner.rho[g] = float(sd[f"rho_{k}"])
        self._pen_scale, self._pen_scale_key = None, None
        self.fused_sums.zero_()


class ConsensusGroup:
    """All blocks trained by this process + the exchange with the other processes."""

    def __init__(self, blocks: Dict[int, object], topo: Topology, alphas: Sequence[float] = DEFAULT_ALPHAS,
                 over_relaxation_coeff: float = 0.5, mu: float = 10.0, tau_inc: float = 2.0, tau_dec: float = 2.0,
                 ops=None, p2p: Optional[bool] = None):
        self.topo = topo
        self.group = topo.group
        self.ops = ops if ops is not None else CudaOps()
        # setup_penalty_parameters (master:326-334): rho_g = alpha_g / N_global
        self.rho = [a / topo.n_global for a in alphas]
        self.relax = 1.0 + over_relaxation_coeff
        self.mu, self.tau_inc, self.tau_dec = mu, tau_inc, tau_dec
        self.last_primal: Optional[List[float]] = None
        self.last_dual: Optional[List[float]] = None
        assert sorted(int(b) for b in blocks) == sorted(topo.plans), "blocks do not match the topology"
        self.blocks: Dict[int, AdmmBlock] = {}
        for b in sorted(topo.plans):
            v = blocks[b]
            self.blocks[b] = v if isinstance(v, AdmmBlock) else AdmmBlock(v, topo.plans[b], _owner=self)
            self.blocks[b].owner = self
        dev = next(iter(self.blocks.values())).x[0].device
        topo.to(dev)
        self.sendbuf = torch.zeros(max(topo.send_rows_total, 1), ROW, dtype=torch.float32, device=dev)
        self.recvbuf = torch.zeros(max(topo.recv_rows_total, 1), ROW, dtype=torch.float32, device=dev)
        self.scratch = self.ops.new_scratch(dev)
        # Exchange transport between GPUs.  Default: the co-owners' receive buffers are mapped into this process
        # (torch symmetric memory: CUDA peer memory over NVLink / NVSwitch) and the pack kernel writes every row
        # straight into the buffer its consumer reads; two device-side barriers order a round.  Where that cannot
        # be set up (one rank, CPU test double, no peer access) — or with p2p=False / DOGS_B200_ADMM_P2P=0 — the
        # rows go through this rank's send buffer and batched NCCL send/recv.
        self.p2p = None
        self.p2p_error = None
        want = (os.environ.get("DOGS_B200_ADMM_P2P", "1") != "0") if p2p is None else bool(p2p)
        if want and self._world() > 1 and topo.exchange == "neighbour" and dev.type == "cuda" and isinstance(self.ops, CudaOps):
            try:
                self._setup_p2p(dev)
            except Exception as ex:  # pylint: disable=W0703
                self.p2p, self.p2p_error = None, repr(ex)
        ok = torch.tensor([1 if self.p2p is not None else 0], dtype=torch.int32, device=dev)
        if self._world() > 1 and dev.type == "cuda":
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)     # all ranks or none
        if int(ok.item()) == 0:
            self.p2p = None
        inv_n = [[1.0 / max(p.n_local * sz, 1) for sz in GROUP_SIZES] + [1.0] * 6 for p in (topo.plans[b] for b in sorted(topo.plans))]
        self._res_scale = torch.tensor(inv_n, dtype=torch.float64, device=dev)

    def _world(self):
        return dist.get_world_size(self.group) if _dist_on(self.group) else 1

    def _setup_p2p(self, dev):
        import torch.distributed._symmetric_memory as symm
        topo = self.topo
        group = self.group if self.group is not None else dist.group.WORLD
        rows = torch.tensor([max(topo.recv_rows_total, 1)], dtype=torch.int64, device=dev)
        dist.all_reduce(rows, op=dist.ReduceOp.MAX, group=self.group)          # symmetric: same size everywhere
        buf = symm.empty(int(rows.item()), ROW, dtype=torch.float32, device=dev)
        hdl = symm.rendezvous(buf, group)
        tables = [None] * topo.world
        dist.all_gather_object(tables, {f"{s_}:{r_}": v for (s_, r_), v in topo.msg_recv.items()}, group=self.group)
        remote = sorted({r for p in topo.plans.values() for (_, r, _, _) in p.messages if r != topo.rank})
        if len(remote) + 1 > 16:
            raise RuntimeError("more than 15 peer GPUs co-own Gaussians with this rank")
        slot_of_rank = {r: 1 + k for k, r in enumerate(remote)}
        for a, p in topo.plans.items():
            sel, row = [], []
            for (b, r, first, n) in p.messages:
                j = torch.arange(n, dtype=torch.int64)
                if r == topo.rank:
                    sel.append(torch.zeros(n, dtype=torch.uint8))
                    row.append(first + j)                              # its place in this rank's own send buffer
                else:
                    sel.append(torch.full((n,), slot_of_rank[r], dtype=torch.uint8))
                    row.append(int(tables[r][f"{a}:{b}"]) + j)          # its place in the receiver's buffer
            p.p2p_sel = (torch.cat(sel) if sel else torch.zeros(0, dtype=torch.uint8)).to(dev)
            p.p2p_row = (torch.cat(row) if row else torch.zeros(0, dtype=torch.int64)).to(torch.int32).to(dev)
        self.recvbuf = buf
        self.p2p = dict(hdl=hdl, dst_ptrs=[self.sendbuf.data_ptr()] + [int(hdl.buffer_ptrs[r]) for r in remote], remote=remote)

    @torch.no_grad()
    def consensus(self, adapt: bool = True):
        """One consensus round: steps (2)-(6) of master_gaussian_trainer.train_iteration (:697-716)."""
        topo = self.topo
        order = sorted(self.blocks)
        if topo.exchange == "dense":
            blk = self.blocks[order[0]]
            self.sendbuf.zero_()
            if blk.plan.S:
                # rows land at their global slot; the other owners' rows are added by the all-reduce
                tmp = torch.empty(blk.plan.S, ROW, dtype=torch.float32, device=self.sendbuf.device)
                dense_plan = _PackView(blk.plan)
                self.ops.pack_rows([t.detach() for t in blk.x], dense_plan, tmp)
                self.sendbuf.index_copy_(0, blk.plan.pack_dst.long(), tmp)
            if self._world() > 1 and topo.S_total > 0:
                dist.all_reduce(self.sendbuf, op=dist.ReduceOp.SUM, group=self.group)
        elif self.p2p is not None:
            hdl = self.p2p["hdl"]
            hdl.barrier(0)       # every co-owner has finished reading last round's rows
            for b in order:
                blk = self.blocks[b]
                self.ops.pack_rows_to([t.detach() for t in blk.x], blk.plan, self.p2p["dst_ptrs"])
            hdl.barrier(0)       # every row of this round is in place
        else:
            for b in order:
                blk = self.blocks[b]
                p = blk.plan
                self.ops.pack_rows([t.detach() for t in blk.x], p, self.sendbuf[p.send_start:p.send_start + p.send_rows])
            self._exchange()
        for b in order:
            blk = self.blocks[b]
            blk.residuals.zero_()
            self.ops.consensus_update([t.detach() for t in blk.x], blk.u, blk.z, blk.zmu, blk.plan,
                                      [self.sendbuf, self.recvbuf], self.relax, self.scratch, blk.residuals)
        # primal_g = sum_k mse(z[idx_k], x_k): every block divides by its own element count
        res = (torch.stack([self.blocks[b].residuals for b in order]) * self._res_scale).sum(0)
        if self._world() > 1:
            dist.all_reduce(res, op=dist.ReduceOp.SUM, group=self.group)
        res = res.tolist()
        primal = res[:6]
        dual = [self.rho[g] * res[6 + g] / (topo.n_global * GROUP_SIZES[g]) for g in range(6)]
        self.last_primal, self.last_dual = primal, dual
        if adapt:  # adapt_penalty_parameters (master:337-377), evaluated identically on every rank
            for g in range(6):
                if primal[g] > self.mu * dual[g]:
                    self.rho[g] *= self.tau_inc
                elif dual[g] > self.mu * primal[g]:
                    self.rho[g] /= self.tau_dec
        return primal, dual

    def _exchange(self):
        """Rows for co-owners on other GPUs: one send per (local block, destination rank), one receive per
        (source rank, sender block), batched.  Sends are posted in (block, rank) order and receives in
        (rank, sender block) order, so the messages between two ranks pair up in block order."""
        if self._world() == 1:
            return
        ops = []
        for b in sorted(self.blocks):
            for r, first, rows in self.blocks[b].plan.sends:
                peer = dist.get_global_rank(self.group, r) if self.group is not None else r
                ops.append(dist.P2POp(dist.isend, self.sendbuf[first:first + rows], peer, self.group))
        for r, first, rows in self.topo.recvs:
            peer = dist.get_global_rank(self.group, r) if self.group is not None else r
            ops.append(dist.P2POp(dist.irecv, self.recvbuf[first:first + rows], peer, self.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()


class _PackView:
    """A dense-mode plan seen by pack_rows as a plain row gather."""

    def __init__(self, plan):
        self.pack_src, self.pack_dst, self.send_rows, self.n_local = plan.pack_src, None, plan.S, plan.n_local


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
