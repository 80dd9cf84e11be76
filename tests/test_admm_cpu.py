"""CPU (no GPU): host logic of the ADMM consensus — topology building, the exchange between blocks of
one process (zero-copy) and between processes (gloo, world_size 2..4), residuals, penalty adaptation,
checkpointing — against tests/golden/admm.npz, which tools/make_golden_admm.py produced by EXECUTING the
reference's own master / slave / GaussianSplatModel methods.  The CUDA kernels are replaced by a torch
test double with the same contract (`ops` injection; the product default is CudaOps, and
tests/test_gpu_aux_kernels.py runs the same golden through the kernels)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dogs_b200 import admm as A

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "admm.npz")
NAMES = A.GROUP_KEYS


class TorchOps:
    """Test double for dogs_b200.admm.CudaOps: same contract, float32 torch arithmetic on CPU tensors
    (IEEE add / divide / multiply, one rounding per operation, like the kernels' __fadd_rn etc.)."""

    def new_scratch(self, device):
        return torch.zeros(1, dtype=torch.uint8, device=device)

    @staticmethod
    def _rows(ts, n):
        return torch.cat([t.reshape(n, sz) for t, sz in zip(ts, A.GROUP_SIZES)], dim=1)

    def pack_rows(self, x, plan, out_rows):
        if plan.send_rows:
            out_rows.copy_(self._rows(x, plan.n_local)[plan.pack_src.long()])

    def consensus_update(self, x, u, z, zmu, plan, sources, relax, scratch, residuals):
        n = plan.n_local
        X, U, Z = self._rows(x, n), self._rows(u, n), self._rows(z, n)
        shared = plan.shared_row >= 0
        z_new, u_new = X.clone(), U.clone()
        cnt = torch.ones(n, dtype=torch.float32)
        if plan.S:
            xs = X[shared]
            s = None
            ent_all = plan.addends.long() & 0xFFFFFFFF
            for j in range(plan.max_addends):
                ent = ent_all[:, j]
                valid = ent != A.NO_ADDEND
                which, row = ent >> 28, ent & 0x0FFFFFFF
                v = xs.clone()
                for k, src in enumerate(sources):
                    m = valid & (which == k)
                    if m.any():
                        v[m] = src[row[m]]
                v[~valid] = 0.0
                s = v if s is None else torch.where(valid[:, None], s + v, s)
            c = plan.count.to(torch.float32)
            cnt[shared] = c
            zs = s / c[:, None]
            z_new[shared] = zs
            u_new[shared] = U[shared] + torch.tensor(relax, dtype=torch.float32) * (xs - zs)
        off = 0
        for g, sz in enumerate(A.GROUP_SIZES):
            sl = slice(off, off + sz)
            residuals[g] += ((z_new[:, sl].double() - X[:, sl].double()) ** 2).sum()
            residuals[6 + g] += (((Z[:, sl].double() - z_new[:, sl].double()) ** 2) / cnt[:, None].double()).sum()
            u[g].copy_(u_new[:, sl].reshape(u[g].shape))
            z[g].copy_(z_new[:, sl].reshape(z[g].shape))
            if zmu is not None:
                zmu[g].copy_((z_new[:, sl] - u_new[:, sl]).reshape(zmu[g].shape))
            off += sz

    def penalty(self, x, u, z, grads, n, coef, sums):
        for g in range(6):
            if x[g] is None:
                continue
            r = (x[g] + u[g]) - z[g] if u is not None else x[g] - z[g]
            sums[g] += (r.double() ** 2).sum()
            if grads is not None:
                grads[g] += coef[g] * r


def load_golden():
    return dict(np.load(GOLDEN))


def run_rounds(G, block_ids, topo, ops):
    """Drives a ConsensusGroup through the golden scenario for the blocks in `block_ids`; returns per round
    what the reference produced: z and u per block, residuals, rho, penalties and their gradients."""
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).clone()
    params = {b: [t(G[f"r0_x_{n}_{b}"]) for n in NAMES] for b in block_ids}
    grp = A.ConsensusGroup(params, topo, ops=ops)
    out = []
    for rnd in range(3):
        rec = {}
        for b in block_ids:
            blk = grp.blocks[b]
            for g, n in enumerate(NAMES):
                blk.x[g].copy_(t(G[f"r{rnd}_x_{n}_{b}"]))       # the seeded "training" between the rounds
            if rnd > 0:
                for x in blk.x:
                    x.grad = torch.zeros_like(x)
                pens = blk.add_penalty_grads_()
                rec[f"pen_{b}"] = pens.numpy().copy()
                rec[f"pengrad_{b}"] = [x.grad.numpy().copy() for x in blk.x]
        rec["rho_before"] = list(grp.rho)
        primal, dual = grp.consensus()
        rec["primal"], rec["dual"], rec["rho_after"] = primal, dual, list(grp.rho)
        for b in block_ids:
            rec[f"z_{b}"] = [a.numpy().copy() for a in grp.blocks[b].z]
            rec[f"u_{b}"] = [a.numpy().copy() for a in grp.blocks[b].u]
            rec[f"zmu_{b}"] = [a.numpy().copy() for a in grp.blocks[b].zmu]
        out.append(rec)
    return out


def check_against_golden(G, block_ids, recs, exact=True):
    for rnd, rec in enumerate(recs):
        assert np.array_equal(np.array(rec["rho_before"]), G[f"r{rnd}_rho_before"])
        assert np.array_equal(np.array(rec["rho_after"]), G[f"r{rnd}_rho_after"]), (rnd, rec["rho_after"])
        # the reference reduces mse_loss in float32; the kernels accumulate in double
        assert np.allclose(rec["primal"], G[f"r{rnd}_primal"], rtol=2e-6, atol=1e-14), rnd
        assert np.allclose(rec["dual"], G[f"r{rnd}_dual"], rtol=2e-6, atol=1e-14), rnd
        for b in block_ids:
            idx = G[f"global_indices_{b}"]
            for g, n in enumerate(NAMES):
                z_ref = G[f"r{rnd}_z_{n}"][idx]
                u_ref = G[f"r{rnd}_u_{n}_{b}"]
                z, u = rec[f"z_{b}"][g].reshape(z_ref.shape), rec[f"u_{b}"][g].reshape(u_ref.shape)
                if exact:      # ascending-block summation + true division: bit for bit the master's average
                    assert np.array_equal(z.view(np.int32), z_ref.view(np.int32)), (rnd, b, n)
                    assert np.array_equal(u.view(np.int32), u_ref.view(np.int32)), (rnd, b, n)
                else:
                    assert np.abs(z - z_ref).max() <= 1e-6 * max(1.0, np.abs(z_ref).max())
                    assert np.abs(u - u_ref).max() <= 4e-6 * max(1.0, np.abs(u_ref).max())
                assert np.array_equal(rec[f"zmu_{b}"][g].reshape(z.shape), z - u)
            if rnd > 0:
                # 0.5 * rho * mse(x + u, z) and its autograd gradient (slave:161-202)
                assert np.allclose(rec[f"pen_{b}"], G[f"r{rnd}_penalty_{b}"], rtol=1e-5, atol=1e-12), (rnd, b)
                for g, n in enumerate(NAMES):
                    ref = G[f"r{rnd}_pengrad_{n}_{b}"]
                    got = rec[f"pengrad_{b}"][g].reshape(ref.shape)
                    # r = x - (z - u) here, (x + u) - z in the reference: one float32 rounding of O(1) values
                    # apart, i.e. ~1e-6 of the largest gradient (the bar for gradients is 1e-4)
                    assert np.abs(got - ref).max() <= 1e-5 * max(np.abs(ref).max(), 1e-30), (rnd, b, n)


def test_golden_all_blocks_in_one_process():
    """8(e) time-sliced protocol: K blocks on one rank exchange through the shared send buffer (no copies)."""
    G = load_golden()
    K = int(G["K"])
    topo = A.build_topology({b: torch.from_numpy(G[f"global_indices_{b}"]) for b in range(K)}, int(G["n_global"]))
    assert topo.recv_rows_total == 0 and topo.send_rows_total > 0
    for b in range(K):
        p = topo.plans[b]
        assert p.S == int((G["visibility_count"][G[f"global_indices_{b}"]] > 1).sum())
        assert p.max_addends == int(G["visibility_count"][G[f"global_indices_{b}"]].max())
        assert not p.sends
    check_against_golden(G, range(K), run_rounds(G, list(range(K)), topo, TorchOps()))


def test_owner_order_is_what_makes_z_exact():
    """Summing the owners in another order changes the last bit for Gaussians with >= 3 owners: the
    table order (ascending block id = the master's plus_gaussians loop) is part of the contract."""
    G = load_golden()
    K = int(G["K"])
    topo = A.build_topology({b: torch.from_numpy(G[f"global_indices_{b}"]) for b in range(K)}, int(G["n_global"]))
    for p in topo.plans.values():            # reverse the addend order of every row
        tab = p.addends.clone()
        for s in range(p.S):
            k = int(p.count[s])
            tab[s, :k] = p.addends[s, :k].flip(0)
        p.addends = tab
    recs = run_rounds(G, list(range(K)), topo, TorchOps())
    check_against_golden(G, range(K), recs[:1], exact=False)
    differs = any(not np.array_equal(recs[1][f"z_{b}"][0].reshape(-1, 3), G["r1_z_xyz"][G[f"global_indices_{b}"]])
                  for b in range(K))
    assert differs


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, assignment, exchange, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        G = load_golden()
        mine = assignment[rank]
        topo = A.build_topology({b: torch.from_numpy(G[f"global_indices_{b}"]) for b in mine}, int(G["n_global"]),
                                exchange=exchange)
        recs = run_rounds(G, mine, topo, TorchOps())
        out[rank] = dict(recs=recs, send=topo.send_rows_total, recv=topo.recv_rows_total,
                         sends={b: list(topo.plans[b].sends) for b in mine}, recvs=list(topo.recvs))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("assignment,exchange", [
    ([[0], [1], [2], [3]], "neighbour"),        # one block per rank (the reference's deployment)
    ([[0, 1], [2, 3]], "neighbour"),            # two time-sliced blocks per rank
    ([[2], [0, 3], [1]], "neighbour"),          # uneven, block ids not in rank order
    ([[0], [1], [2], [3]], "dense"),            # literal all-reduce of the dense shared buffer
])
def test_multi_rank_consensus_over_gloo_matches_the_reference(assignment, exchange):
    world = len(assignment)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), assignment, exchange, out), nprocs=world, join=True)
    G = load_golden()
    for rank, mine in enumerate(assignment):
        # dense: the all-reduce adds the owners in gloo's order, not the master's => last-bit differences
        check_against_golden(G, mine, out[rank]["recs"], exact=(exchange == "neighbour"))
        for rnd in range(3):                    # residuals and rho identical on every rank (no master)
            assert out[rank]["recs"][rnd]["primal"] == out[0]["recs"][rnd]["primal"]
            assert out[rank]["recs"][rnd]["rho_after"] == out[0]["recs"][rnd]["rho_after"]
    if exchange == "neighbour":
        # every row sent to another rank is received there
        sent = sum(rows for r in range(world) for b in assignment[r] for _, _, rows in out[r]["sends"][b])
        recvd = sum(rows for r in range(world) for _, _, rows in out[r]["recvs"])
        assert sent == recvd and sent > 0


def test_state_dict_round_trip_with_reference_keys():
    """checkpoints use the reference slave's keys (u_xyz .. rho_o, slave:209-263) and can be loaded;
    after a reload the penalty and the next consensus round are identical."""
    G = load_golden()
    K = int(G["K"])
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).clone()
    mk = lambda: A.build_topology({b: torch.from_numpy(G[f"global_indices_{b}"]) for b in range(K)}, int(G["n_global"]))
    grp = A.ConsensusGroup({b: [t(G[f"r1_x_{n}_{b}"]) for n in NAMES] for b in range(K)}, mk(), ops=TorchOps())
    grp.consensus()
    sds = {b: grp.blocks[b].state_dict() for b in range(K)}
    sd = sds[1]
    assert {f"u_{k}" for k in NAMES} | {f"rho_{k}" for k in NAMES} | {"admm_enabled"} <= set(sd)
    assert sd["u_fdc"].shape == G["r1_x_fdc_1"].shape and sd["u_fr"].shape == G["r1_x_fr_1"].shape
    grp2 = A.ConsensusGroup({b: [t(G[f"r1_x_{n}_{b}"]) for n in NAMES] for b in range(K)}, mk(), ops=TorchOps())
    for b in range(K):
        grp2.blocks[b].load_state_dict(sds[b])
    assert grp2.rho == grp.rho
    for g_ in (grp, grp2):
        for b in range(K):
            for gi, n in enumerate(NAMES):
                g_.blocks[b].x[gi].copy_(t(G[f"r2_x_{n}_{b}"]))
                g_.blocks[b].x[gi].grad = torch.zeros_like(g_.blocks[b].x[gi])
    for b in range(K):
        assert torch.equal(grp.blocks[b].add_penalty_grads_(), grp2.blocks[b].add_penalty_grads_())
        for a, c in zip(grp.blocks[b].x, grp2.blocks[b].x):
            assert torch.equal(a.grad, c.grad)
    assert grp.consensus() == grp2.consensus()
    for b in range(K):
        for a, c in zip(grp.blocks[b].u + grp.blocks[b].z, grp2.blocks[b].u + grp2.blocks[b].z):
            assert torch.equal(a, c)
    # a reference checkpoint has no z_*: the consensus copy restarts at z = x
    ref_sd = {k: v for k, v in sd.items() if not k.startswith("z_")}
    grp2.blocks[1].load_state_dict(ref_sd)
    for x, z, u, zmu in zip(grp2.blocks[1].x, grp2.blocks[1].z, grp2.blocks[1].u, grp2.blocks[1].zmu):
        assert torch.equal(z, x.detach()) and torch.equal(zmu, z - u)
    with pytest.raises(ValueError, match="does not match the block"):
        grp2.blocks[0].load_state_dict(sd)


def test_cuda_ops_is_the_default_and_needs_the_library():
    ops = A.CudaOps()       # loads libdogs_b200.so (raises ImportError if it was not built)
    assert ops.lib.dogs_abi_version() == 3
    with pytest.raises(RuntimeError, match="contiguous float32 CUDA tensors"):
        ops._t6([torch.zeros(3, 3)] * 6)
