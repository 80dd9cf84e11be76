"""GPU parity of fused-SSIM, simple-knn, sparse Adam and the ADMM kernels against the unmodified
reference build (oracle/_ref), the CPU oracle and the golden fixtures."""
import os

import numpy as np
import pytest
import torch

import refutil as U

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _ref():
    if not U.ref_available():
        pytest.skip("oracle/_ref not present")
    return U.ref_modules()


@pytest.mark.parametrize("shape", [(1, 3, 840, 1297), (2, 3, 67, 45), (1, 1, 5, 7), (1, 3, 32, 32)])
@pytest.mark.parametrize("padding", ["same", "valid"])
def test_fused_ssim_vs_reference(shape, padding):
    _, rssim, _ = _ref()
    import dogs_b200.fused_ssim as mine
    if padding == "valid" and min(shape[2:]) <= 10:
        pytest.skip("valid padding needs > 10 px")
    g = torch.Generator(device="cpu").manual_seed(0)
    a = torch.rand(*shape, generator=g).to(DEV)
    b = (a.cpu() * 0.6 + 0.4 * torch.rand(*shape, generator=g)).to(DEV)
    a1, a2 = a.clone().requires_grad_(True), a.clone().requires_grad_(True)
    l1 = rssim.fused_ssim(a1, b, padding=padding)
    l2 = mine.fused_ssim(a2, b, padding=padding)
    l1.backward()
    l2.backward()
    assert abs(l1.item() - l2.item()) <= 1e-6            # fp tolerance of the fused-SSIM value
    assert U.rel_to_max(a2.grad, a1.grad) <= 1e-5
    assert abs(rssim.fused_ssim(a, b, train=False).item() - mine.fused_ssim(a, b, train=False).item()) <= 1e-6


def test_fused_ssim_golden_and_oracle():
    import dogs_b200.fused_ssim as mine
    from oracle import oracle as O
    g = np.load(os.path.join(GOLD, "ssim.npz"))
    a, b = torch.from_numpy(g["img1"]).to(DEV), torch.from_numpy(g["img2"]).to(DEV)
    m, d1, d2, d3 = mine.fusedssim(0.01 ** 2, 0.03 ** 2, a, b, True)
    assert np.abs(m.cpu().numpy() - g["ssim_map"]).max() <= 2e-6
    assert np.abs(d1.cpu().numpy() - g["dm_dmu1"]).max() <= 1e-5 * np.abs(g["dm_dmu1"]).max()
    gi = mine.fusedssim_backward(0.01 ** 2, 0.03 ** 2, a, b, torch.from_numpy(g["dL_dmap"]).to(DEV), d1, d2, d3)
    assert np.abs(gi.cpu().numpy() - g["dL_dimg1"]).max() <= 1e-5 * np.abs(g["dL_dimg1"]).max()
    om = O.ssim_fwd(g["img1"], g["img2"])[0]
    assert np.abs(om - m.cpu().numpy()).max() <= 2e-6


@pytest.mark.parametrize("P", [1_100_000, 100_000, 4097, 1024, 5, 3])
def test_simple_knn_bit_exact(P):
    # (1.1 M points: Morton sort of more than 2^20 keys = the three-kernel radix passes; the smaller sizes run the
    #  chained single-kernel passes, csrc/radix_sort.cu)
    _, _, rknn = _ref()
    from dogs_b200.simple_knn._C import distCUDA2
    g = torch.Generator(device="cpu").manual_seed(P)
    pts = (torch.randn(P, 3, generator=g) * torch.tensor([3.0, 1.0, 0.2])).to(DEV)
    if P > 1000:
        pts[17] = pts[23]                                  # duplicate point => a zero distance
    d1 = rknn.distCUDA2(pts)
    d2 = distCUDA2(pts)
    torch.cuda.synchronize()
    assert torch.equal(U.bits(d1), U.bits(d2))


def test_simple_knn_golden_and_oracle():
    from dogs_b200.simple_knn._C import distCUDA2
    from oracle import oracle as O
    g = np.load(os.path.join(GOLD, "knn.npz"))
    d = distCUDA2(torch.from_numpy(g["points"]).to(DEV)).cpu().numpy()
    assert np.array_equal(d.view(np.int32), g["dist2"].view(np.int32))
    assert np.array_equal(O.knn_dist2(g["points"]).view(np.int32), d.view(np.int32))


@pytest.mark.parametrize("N,M", [(10000, 45), (10000, 3), (777, 1), (1001, 4), (64, 9)])
def test_sparse_adam_bit_exact(N, M):
    dgr, _, _ = _ref()
    import dogs_b200.diff_gaussian_rasterization as mine
    g = torch.Generator(device="cpu").manual_seed(N + M)
    p = torch.randn(N, M, generator=g).to(DEV)
    gr = torch.randn(N, M, generator=g).to(DEV)
    m = (torch.randn(N, M, generator=g) * 0.1).to(DEV)
    v = (torch.rand(N, M, generator=g) * 0.01).to(DEV)
    vis = (torch.rand(N, generator=g) > 0.4).to(DEV)
    pa, ma, va = p.clone(), m.clone(), v.clone()
    pb, mb, vb = p.clone(), m.clone(), v.clone()
    dgr._C.adamUpdate(pa, gr, ma, va, vis, 1e-3, 0.9, 0.999, 1e-15, N, M)
    mine.adamUpdate(pb, gr, mb, vb, vis, 1e-3, 0.9, 0.999, 1e-15, N, M)
    torch.cuda.synchronize()
    for x, y in ((pa, pb), (ma, mb), (va, vb)):
        assert torch.equal(U.bits(x), U.bits(y))
    assert torch.equal(pb[~vis], p[~vis])


def test_sparse_adam_optimizer_class():
    """SparseGaussianAdam.step(visibility, N) (reference __init__.py:303-332) end to end."""
    dgr, _, _ = _ref()
    import dogs_b200.diff_gaussian_rasterization as mine
    g = torch.Generator(device="cpu").manual_seed(0)
    N = 300
    res = []
    for mod in (dgr, mine):
        w = torch.randn(N, 15, 3, generator=torch.Generator(device="cpu").manual_seed(1)).to(DEV).requires_grad_(True)
        opt = mod.SparseGaussianAdam([{"params": [w], "lr": 1e-2, "name": "f_rest"}], lr=0.0, eps=1e-15)
        for it in range(3):
            w.grad = torch.randn(N, 15, 3, generator=torch.Generator(device="cpu").manual_seed(10 + it)).to(DEV)
            vis = (torch.rand(N, generator=torch.Generator(device="cpu").manual_seed(20 + it)) > 0.3).to(DEV)
            opt.step(vis, N)
        res.append(w.detach().clone())
    assert torch.equal(U.bits(res[0]), U.bits(res[1]))


def test_multi_tensor_sparse_adam_matches_reference_optimizer():
    """one launch for the six parameter tensors (dgr_sparse_adam_multi) against the reference's
    SparseGaussianAdam (six adamUpdate launches), bit for bit, with the visibility given as a bool mask and as
    the rasterizer's int32 radii; N not a multiple of 4 so every tensor has a ragged tail."""
    dgr, _, _ = _ref()
    import dogs_b200.diff_gaussian_rasterization as mine
    from dogs_b200 import _lib
    N = 10_003
    shapes = {"xyz": (N, 3), "f_dc": (N, 1, 3), "f_rest": (N, 15, 3), "opacity": (N, 1), "scaling": (N, 3), "quaternion": (N, 4)}
    lrs = {"xyz": 1.6e-4, "f_dc": 2.5e-3, "f_rest": 1.25e-4, "opacity": 0.05, "scaling": 5e-3, "quaternion": 1e-3}
    res = {}
    for tag, mod, use_radii in (("ref", dgr, False), ("mask", mine, False), ("radii", mine, True)):
        g = torch.Generator(device="cpu").manual_seed(1)
        ps = {k: torch.randn(*s, generator=g).to(DEV).requires_grad_(True) for k, s in shapes.items()}
        opt = mod.SparseGaussianAdam([{"params": [ps[k]], "lr": lrs[k], "name": k} for k in shapes], lr=0.0, eps=1e-15)
        n0 = _lib.lib().dogs_launch_count()
        for it in range(3):
            gg = torch.Generator(device="cpu").manual_seed(10 + it)
            for k in shapes:
                ps[k].grad = torch.randn(*shapes[k], generator=gg).to(DEV)
            radii = (torch.rand(N, generator=gg) * 20 - 6).to(torch.int32).clamp_min(0).to(DEV)
            opt.step(radii if use_radii else radii > 0, N)
        if mod is mine:
            assert _lib.lib().dogs_launch_count() - n0 == 3          # one launch per step for all six tensors
        res[tag] = {k: v.detach().clone() for k, v in ps.items()}
        res[tag + "_m"] = {k: opt.state[ps[k]]["exp_avg"].clone() for k in shapes}
    for k in shapes:
        for tag in ("mask", "radii"):
            assert torch.equal(U.bits(res[tag][k]), U.bits(res["ref"][k])), (tag, k)
            assert torch.equal(U.bits(res[tag + "_m"][k]), U.bits(res["ref_m"][k])), (tag, k)


def _virtual_plan(P, A):
    """One block whose Gaussians all have a second (virtual) owner: enough for the per-iteration penalty."""
    plan = A.BlockPlan(0, P, P)
    plan.shared_row = torch.arange(P, dtype=torch.int32, device=DEV)
    plan.S, plan.max_addends = P, 1
    plan.addends = torch.full((P, 1), A.SELF_ADDEND - (1 << 32), dtype=torch.int32, device=DEV)   # uint32 bit pattern
    plan.count = torch.ones(P, dtype=torch.uint8, device=DEV)
    plan.pack_src = torch.zeros(0, dtype=torch.int32, device=DEV)
    topo = A.Topology()
    topo.n_global = P
    topo.plans[0] = plan
    return topo


def test_admm_consensus_kernels_reproduce_the_reference_golden():
    """admm_pack_rows -> (zero-copy exchange between the 4 blocks of one GPU) -> admm_consensus_update against
    tests/golden/admm.npz, i.e. against the reference's own master/slave methods executed on the same inputs:
    z and u BIT-EXACT (owners added in block order, IEEE division by the count, torch's rounding of
    u += 1.5 * (x - z)), residuals to 2e-6, rho identical, penalty and its gradients to 1e-5."""
    from dogs_b200 import admm as A
    import test_admm_cpu as T
    G = T.load_golden()
    K = int(G["K"])
    topo = A.build_topology({b: torch.from_numpy(G[f"global_indices_{b}"]).to(DEV) for b in range(K)}, int(G["n_global"]))
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    grp = A.ConsensusGroup({b: [t(G[f"r0_x_{n}_{b}"]) for n in T.NAMES] for b in range(K)}, topo)
    assert isinstance(grp.ops, A.CudaOps)
    from dogs_b200 import _lib
    n0 = _lib.lib().dogs_launch_count()
    recs = []
    for rnd in range(3):
        rec = {}
        for b in range(K):
            blk = grp.blocks[b]
            for g, n in enumerate(T.NAMES):
                blk.x[g].copy_(t(G[f"r{rnd}_x_{n}_{b}"]))
            if rnd > 0:
                for x in blk.x:
                    x.grad = torch.zeros_like(x)
                rec[f"pen_{b}"] = blk.add_penalty_grads_().cpu().numpy()
                rec[f"pengrad_{b}"] = [x.grad.cpu().numpy() for x in blk.x]
        rec["rho_before"] = list(grp.rho)
        primal, dual = grp.consensus()
        rec["primal"], rec["dual"], rec["rho_after"] = primal, dual, list(grp.rho)
        for b in range(K):
            rec[f"z_{b}"] = [a.cpu().numpy() for a in grp.blocks[b].z]
            rec[f"u_{b}"] = [a.cpu().numpy() for a in grp.blocks[b].u]
            rec[f"zmu_{b}"] = [a.cpu().numpy() for a in grp.blocks[b].zmu]
        recs.append(rec)
    T.check_against_golden(G, range(K), recs)
    # 2 launches per block and round (pack, consensus_update) + 1 penalty launch per block in rounds 2 and 3
    assert _lib.lib().dogs_launch_count() - n0 == 3 * K * 2 + 2 * K


def test_admm_consensus_kernels_large_random_vs_torch():
    """Sizes where every CTA of the two kernels has work (ragged last warp, several addends): against the
    same arithmetic in torch on the GPU (float32 adds in owner order, true division)."""
    from dogs_b200 import admm as A
    rng = np.random.default_rng(5)
    n_global, K = 300_007, 5
    member = rng.random((n_global, K)) < 0.35
    member[member.sum(1) == 0, 2] = True
    gidx = [torch.from_numpy(np.nonzero(member[:, k])[0]).to(DEV) for k in range(K)]
    topo = A.build_topology({k: gidx[k] for k in range(K)}, n_global)
    g = torch.Generator(device="cpu").manual_seed(1)
    xs = [[torch.randn(len(gidx[k]), sz, generator=g).to(DEV) for sz in A.GROUP_SIZES] for k in range(K)]
    grp = A.ConsensusGroup({k: [a.clone() for a in xs[k]] for k in range(K)}, topo)
    u0 = [[torch.randn(a.shape, generator=g).to(DEV) * 0.1 for a in xs[k]] for k in range(K)]
    for k in range(K):
        for a, b in zip(grp.blocks[k].u, u0[k]):
            a.copy_(b)
    primal, dual = grp.consensus(adapt=False)
    count = torch.from_numpy(member.sum(1)).to(DEV)
    for gi, sz in enumerate(A.GROUP_SIZES):
        zsum = torch.zeros(n_global, sz, device=DEV)
        for k in range(K):                      # ascending block order, one float32 add per owner
            zsum[gidx[k]] += xs[k][gi]
        z = zsum / count[:, None].float()
        for k in range(K):
            zl = z[gidx[k]]
            shared = (count[gidx[k]] > 1)[:, None]
            want_z = torch.where(shared, zl, xs[k][gi])
            want_u = torch.where(shared, u0[k][gi] + torch.tensor(1.5, device=DEV) * (xs[k][gi] - zl), u0[k][gi])
            assert torch.equal(grp.blocks[k].z[gi], want_z), (gi, k)
            assert torch.equal(grp.blocks[k].u[gi], want_u), (gi, k)
            assert torch.equal(grp.blocks[k].zmu[gi], want_z - want_u)
    ref_primal = [sum(float(((grp.blocks[k].z[gi].double() - xs[k][gi].double()) ** 2).mean()) for k in range(K)) for gi in range(6)]
    assert np.allclose(primal, ref_primal, rtol=1e-9)


def test_fused_admm_penalty_in_rasterizer_backward_matches_standalone():
    """DgrAdmmPenalty: the rasterizer backward adds the xyz / f_dc / f_rest penalty gradients itself;
    the result must equal rasterizer gradient + standalone admm_penalty (all six groups)."""
    from dogs_b200 import admm as A, scenes
    import dogs_b200.diff_gaussian_rasterization as dgr
    sc = scenes.make_scene("tiny", 11, DEV)
    P = sc["means3D"].shape[0]
    g = torch.Generator(device="cpu").manual_seed(3)
    w = torch.randn(3, sc["H"], sc["W"], generator=g).to(DEV)

    def leaves():
        d = {"means3D": sc["means3D"].clone(), "dc": sc["dc"].reshape(P, 3).clone(), "sh": sc["sh"].reshape(P, 45).clone(),
             "scales": sc["scales"].clone(), "rotations": sc["rotations"].clone(), "opacities": sc["opacities"].clone()}
        for v in d.values():
            v.requires_grad_(True)
        return d

    def run(fused, all_groups=False):
        lv = leaves()
        x = [lv[k] for k in ("means3D", "dc", "sh", "scales", "rotations", "opacities")]
        blk = A.AdmmBlock(x, _virtual_plan(P, A), alphas=(3e3, 2e3, 5e3, 1e3, 7e3, 9e2))
        gg = torch.Generator(device="cpu").manual_seed(5)
        for t in blk.zmu:                                   # pretend a consensus round happened
            t.add_(0.05 * torch.randn(t.shape, generator=gg).to(DEV))
        blk.fuse_into_rasterizer(fused, all_groups=all_groups)
        views = dict(lv)
        views["dc"] = lv["dc"].view(P, 1, 3)
        views["sh"] = lv["sh"].view(P, 15, 3)
        # the block travels with the settings of this one render (no module-level state)
        img, _, _, _ = scenes.render(dgr, sc, views, admm=blk if fused else None)
        (img * w).sum().backward()
        pens = blk.add_penalty_grads_()
        torch.cuda.synchronize()
        return {k: v.grad.clone() for k, v in lv.items()}, pens.clone()

    g0, p0 = run(False)
    g1, p1 = run(True)
    assert torch.allclose(p0, p1, rtol=1e-6, atol=1e-15)
    for k in g0:
        assert U.rel_to_max(g1[k], g0[k]) <= 2e-6, k
    # scaling / quaternion / opacity can only be fused when the rasterizer is fed the raw consensus tensors
    # (tests/test_gpu_raw_activations.py covers that path): refusing is better than a silently wrong penalty
    with pytest.raises(RuntimeError, match="needs raw_activations=True"):
        run(True, all_groups=True)


def test_importance_count_render_vs_oracle():
    """f_count mode (LightGaussian statistics): counts are exact integers, so they must equal the
    oracle's wherever its n_contrib agrees (libm-vs-libdevice expf flips aside), and the image is the
    ordinary render."""
    from dogs_b200 import scenes
    import dogs_b200.diff_gaussian_rasterization as dgr
    from oracle import oracle as O
    sc = scenes.make_scene("tiny", 12, DEV)
    rs = scenes.raster_settings(dgr, sc)._replace(f_count=True)
    rast = dgr.GaussianRasterizer(rs)
    cnt, score, img, radii = rast(means3D=sc["means3D"], means2D=torch.zeros_like(sc["means3D"]), opacities=sc["opacities"],
                                  dc=sc["dc"], shs=sc["sh"], scales=sc["scales"], rotations=sc["rotations"])
    ref_img = scenes.render(dgr, sc)[0]
    assert torch.equal(img, ref_img)
    s = {k: (v.cpu().numpy() if isinstance(v, torch.Tensor) else v) for k, v in sc.items()}
    pre = O.preprocess(s)
    bins = O.binning(s, pre)
    o = O.render_fwd(s, pre, bins, count=True)
    c = cnt.cpu().numpy()
    assert int(c.sum()) > 0 and c.dtype == np.int32
    assert int(np.abs(c.astype(np.int64) - o["gaussians_count"]).sum()) <= 8      # expf threshold flips only
    same = c == o["gaussians_count"]
    assert np.allclose(score.cpu().numpy()[same], o["important_score"][same], rtol=1e-6)
    assert int((c[radii.cpu().numpy() == 0] != 0).sum()) == 0
    # the call style of the reference's count_render(): keywords, no `antialiasing`, f_count=True
    rs2 = dgr.GaussianRasterizationSettings(image_height=sc["H"], image_width=sc["W"], tanfovx=sc["tanfovx"],
                                            tanfovy=sc["tanfovy"], bg=torch.zeros(3, device=DEV), scale_modifier=1.0,
                                            depth_threshold=0.0, viewmatrix=sc["viewmatrix"], projmatrix=sc["projmatrix"],
                                            sh_degree=3, campos=sc["campos"], prefiltered=False, debug=False, f_count=True)
    out = dgr.GaussianRasterizer(rs2)(means3D=sc["means3D"], means2D=None, opacities=sc["opacities"],
                                      shs=torch.cat([sc["dc"], sc["sh"]], 1), scales=sc["scales"], rotations=sc["rotations"])
    assert len(out) == 4 and torch.equal(out[0], cnt)


def test_compaction_equals_boolean_indexing():
    """dogs_b200.compaction.compact == [t[mask] for t in tensors] (prune bookkeeping)."""
    from dogs_b200.compaction import compact
    g = torch.Generator(device="cpu").manual_seed(0)
    for P in (1, 7, 100_003):
        ts = [torch.randn(P, 3, generator=g).to(DEV), torch.randn(P, 15, 3, generator=g).to(DEV),
              torch.randn(P, 1, generator=g).to(DEV), torch.randn(P, 4, generator=g).to(DEV),
              torch.randint(0, 1000, (P,), generator=g, dtype=torch.int32).to(DEV)]
        for frac in (0.0, 0.37, 1.0):
            keep = (torch.rand(P, generator=g) < frac).to(DEV)
            out = compact(ts, keep)
            for a, t in zip(out, ts):
                assert a.shape == t[keep].shape and torch.equal(a, t[keep])
