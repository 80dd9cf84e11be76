"""CPU (no GPU): host logic of the ADMM consensus — plan building, the multi-rank exchange
(world_size 2 over gloo), residuals and penalty adaptation — driven with the oracle's numpy
arithmetic in place of the CUDA kernels (`ops` injection; the product default is CudaOps)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dogs_b200 import admm as A
from oracle import oracle as O


class OracleOps:
    """Test double for dogs_b200.admm.CudaOps: same contract, numpy on CPU tensors."""

    def pack(self, x, plan, sendbuf):
        buf = sendbuf.view(-1, A.ROW)
        buf.zero_()
        off = 0
        for g, sz in enumerate(A.GROUP_SIZES):
            buf[plan.slot, off:off + sz] = x[g].view(plan.n_local, sz)[plan.local_idx]
            off += sz

    def unpack_update(self, x, u, z, plan, recvbuf, relax, residuals, zmu=None):
        buf = recvbuf.view(-1, A.ROW)
        off = 0
        shared = plan.slot_of_local >= 0
        for g, sz in enumerate(A.GROUP_SIZES):
            xg, ug, zg = (t.view(plan.n_local, sz) for t in (x[g], u[g], z[g]))
            z_new = xg.clone()
            z_new[shared] = buf[plan.slot_of_local[shared], off:off + sz] * plan.inv_count[shared, None]
            ug[shared] += relax * (xg[shared] - z_new[shared])
            residuals[g] += ((z_new - xg).double() ** 2).sum()
            residuals[6 + g] += (plan.inv_count[:, None].double() * (zg - z_new).double() ** 2).sum()
            zg.copy_(z_new)
            if zmu is not None:
                zmu[g].view(plan.n_local, sz).copy_(z_new - ug)
            off += sz

    def penalty(self, x, u, z, grads, n, coef, sums):
        for g in range(6):
            if x[g] is None:
                continue
            r = (x[g] + u[g]) - z[g] if u is not None else x[g] - z[g]
            sums[g] += (r.double() ** 2).sum()
            if grads is not None:
                grads[g] += coef[g] * r


def _problem(nblocks, n_global=400, seed=0, shuffle=False):
    rng = np.random.default_rng(seed)
    gidx, xs = [], []
    owned = np.zeros(n_global, bool)
    for k in range(nblocks):
        idx = np.sort(rng.choice(n_global, size=int(n_global * 0.6), replace=False))
        gidx.append(idx)
        owned[idx] = True
    gidx[0] = np.sort(np.union1d(gidx[0], np.nonzero(~owned)[0]))
    if shuffle:          # blocks need not list their Gaussians in ascending global order
        gidx = [rng.permutation(idx) for idx in gidx]
    for idx in gidx:
        xs.append([rng.standard_normal((len(idx), sz)).astype(np.float32) for sz in O.GROUP_SIZES])
    return gidx, xs


def test_plan_matches_bincount_semantics():
    gidx, _ = _problem(3)
    count = np.zeros(400, np.int32)
    for idx in gidx:
        count[idx] += 1
    plan = A.ConsensusPlan(torch.from_numpy(gidx[1]), torch.from_numpy(count))
    assert plan.n_local == len(gidx[1]) and plan.S_total == int((count > 1).sum())
    shared_local = count[gidx[1]] > 1
    assert plan.S == int(shared_local.sum())
    assert np.array_equal(plan.local_idx.numpy(), np.nonzero(shared_local)[0])
    assert np.allclose(plan.inv_count.numpy(), 1.0 / count[gidx[1]])
    slots = plan.slot.numpy()
    assert len(np.unique(slots)) == len(slots) and slots.max() < plan.S_total
    # slot numbering is the same on every block: global id -> slot
    other = A.ConsensusPlan(torch.from_numpy(gidx[2]), torch.from_numpy(count))
    common = np.intersect1d(gidx[1], gidx[2])
    common = common[count[common] > 1]
    s1 = plan.slot_of_local.numpy()[np.searchsorted(gidx[1], common)]
    s2 = other.slot_of_local.numpy()[np.searchsorted(gidx[2], common)]
    assert np.array_equal(s1, s2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, gidx, xs, out, exchange):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n_global = 400
        plan = A.build_plan(torch.from_numpy(gidx[rank]), n_global, exchange=exchange)
        assert (plan.peers is not None) == (exchange == "neighbour")
        blk = A.AdmmBlock([torch.from_numpy(a.copy()) for a in xs[rank]], plan, ops=OracleOps())
        rho0 = list(blk.rho)
        primal, dual = blk.consensus()
        # second round after a local "training" perturbation of x
        for t in blk.x:
            t.add_(0.01 * (rank + 1))
        primal2, dual2 = blk.consensus()
        for t in blk.x:
            t.grad = torch.zeros_like(t)
        pens = blk.add_penalty_grads_()
        out[rank] = dict(z=[t.numpy().copy() for t in blk.z], u=[t.numpy().copy() for t in blk.u], primal=primal,
                         dual=dual, primal2=primal2, dual2=dual2, rho0=rho0, rho=list(blk.rho),
                         pens=pens.numpy().copy(), grads=[t.grad.numpy().copy() for t in blk.x])
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,exchange,shuffle", [(2, "neighbour", False), (3, "neighbour", False),
                                                    (3, "neighbour", True), (3, "dense", False)])
def test_multi_rank_consensus_over_gloo_matches_reference_arithmetic(world, exchange, shuffle):
    """world 3: Gaussians owned by three blocks, i.e. rows that receive from two peers.  "neighbour" is the
    point-to-point exchange with the co-owners only, "dense" the all-reduce over all shared Gaussians."""
    gidx, xs = _problem(world, shuffle=shuffle)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), gidx, xs, out, exchange), nprocs=world, join=True)
    n_global = 400
    # reference arithmetic (oracle numpy restatement of master:538-555, slave:99-121, master:396-456)
    z, count = O.admm_consensus(xs, gidx, n_global)
    rho = [a / n_global for a in A.DEFAULT_ALPHAS]
    z_prev = [np.zeros((n_global, sz), np.float32) for sz in O.GROUP_SIZES]
    for k in range(world):               # before the first round z_local == x (fusion hand-off)
        for g in range(6):
            z_prev[g][gidx[k]] = xs[k][g]
    primal_ref, _ = O.admm_residuals(xs, gidx, z, z, rho)
    u1 = [O.admm_dual_update(xs[k], [np.zeros_like(a) for a in xs[k]], [zz[gidx[k]] for zz in z], 1.5) for k in range(world)]
    for k in range(world):
        assert np.allclose(out[k]["primal"], primal_ref, rtol=1e-5)
        assert out[0]["primal"] == out[k]["primal"] and out[0]["dual"] == out[k]["dual"]   # identical on all ranks
        assert out[0]["rho"] == out[k]["rho"]
    # round 2
    xs2 = [[a + np.float32(0.01 * (k + 1)) for a in xs[k]] for k in range(world)]
    z2, _ = O.admm_consensus(xs2, gidx, n_global)
    rho1 = O.admm_adapt(rho, out[0]["primal"], out[0]["dual"])
    primal2_ref, dual2_ref = O.admm_residuals(xs2, gidx, z2, z, rho1)
    for k in range(world):
        u2 = O.admm_dual_update(xs2[k], u1[k], [zz[gidx[k]] for zz in z2], 1.5)
        shared = count[gidx[k]] > 1
        for g in range(6):
            assert np.abs(out[k]["z"][g] - z2[g][gidx[k]]).max() <= 1e-6
            # the reference updates u for every Gaussian; for count==1, z==x so u stays 0 in both
            assert np.abs(out[k]["u"][g] - u2[g]).max() <= 1e-5
        assert np.allclose(out[k]["primal2"], primal2_ref, rtol=1e-5)
        assert np.allclose(out[k]["dual2"], dual2_ref, rtol=1e-4, atol=1e-12)
        rho2 = O.admm_adapt(rho1, primal2_ref, dual2_ref)
        assert np.allclose(out[k]["rho"], rho2, rtol=1e-12)
        pens_ref, grads_ref = O.admm_penalty(xs2[k], u2, [zz[gidx[k]] for zz in z2], rho2)
        assert np.allclose(out[k]["pens"], pens_ref, rtol=1e-5, atol=1e-15)
        for g in range(6):
            assert np.abs(out[k]["grads"][g] - grads_ref[g]).max() <= 1e-6 * max(1.0, float(np.abs(grads_ref[g]).max()))


def test_cuda_ops_is_the_default_and_needs_the_library():
    ops = A.CudaOps()       # loads libdogs_b200.so (raises ImportError if it was not built)
    assert ops.lib.dogs_abi_version() == 1
    with pytest.raises(RuntimeError, match="contiguous float32 CUDA tensors"):
        ops._t6([torch.zeros(3, 3)] * 6)
