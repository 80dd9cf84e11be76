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


@pytest.mark.parametrize("P", [100_000, 4097, 1024, 5, 3])
def test_simple_knn_bit_exact(P):
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


def _admm_problem(nblocks=3, n_global=5000, seed=0):
    rng = np.random.default_rng(seed)
    from oracle import oracle as O
    gidx, xs = [], []
    for k in range(nblocks):
        idx = np.sort(rng.choice(n_global, size=int(n_global * 0.55), replace=False))
        gidx.append(idx)
        xs.append([rng.standard_normal((len(idx), sz)).astype(np.float32) for sz in O.GROUP_SIZES])
    # make sure every global Gaussian is owned by someone
    owned = np.zeros(n_global, bool)
    for idx in gidx:
        owned[idx] = True
    missing = np.nonzero(~owned)[0]
    if len(missing):
        gidx[0] = np.sort(np.concatenate([gidx[0], missing]))
        xs[0] = [rng.standard_normal((len(gidx[0]), sz)).astype(np.float32) for sz in O.GROUP_SIZES]
    return gidx, xs


def test_admm_consensus_kernels_vs_oracle():
    """pack -> (sum over blocks) -> unpack_update on one GPU for 3 blocks == the reference's torch
    arithmetic restated in numpy (oracle.admm_*)."""
    from oracle import oracle as O
    from dogs_b200 import admm as A
    gidx, xs = _admm_problem()
    n_global = 5000
    z_ref, count = O.admm_consensus(xs, gidx, n_global)
    count_t = torch.from_numpy(count.astype(np.int32)).to(DEV)
    blocks = []
    for idx, x in zip(gidx, xs):
        plan = A.ConsensusPlan(torch.from_numpy(idx).to(DEV), count_t).to(DEV)
        blocks.append(A.AdmmBlock([torch.from_numpy(a).to(DEV) for a in x], plan))
    # emulate the all-reduce on one device: sum the packed buffers of all blocks
    total = torch.zeros_like(blocks[0].buf)
    for b in blocks:
        b.ops.pack(b.x, b.plan, b.buf)
        total += b.buf
    z_prev = [[z.clone() for z in b.z] for b in blocks]
    prim = np.zeros(6)
    dual = np.zeros(6)
    for b in blocks:
        b.residuals.zero_()
        b.ops.unpack_update(b.x, b.u, b.z, b.plan, total, b.relax, b.residuals, b.zmu)
        r = b.residuals.cpu().numpy()
        prim += r[:6] / (b.plan.n_local * np.array(O.GROUP_SIZES))
        dual += r[6:]
    torch.cuda.synchronize()
    for k, (b, idx) in enumerate(zip(blocks, gidx)):
        u_ref = O.admm_dual_update(xs[k], [np.zeros_like(a) for a in xs[k]], [z[idx] for z in z_ref], 1.5)
        for g in range(6):
            assert np.abs(b.z[g].cpu().numpy() - z_ref[g][idx]).max() <= 1e-6
            assert np.abs(b.u[g].cpu().numpy() - u_ref[g]).max() <= 1e-5
    rho = [1.0] * 6
    zp = [np.zeros((n_global, sz), np.float32) for sz in O.GROUP_SIZES]
    for k, idx in enumerate(gidx):
        for g in range(6):
            zp[g][idx] = z_prev[k][g].cpu().numpy()       # z_prev = x of the owning blocks (any copy)
    p_ref, _ = O.admm_residuals(xs, gidx, z_ref, z_ref, rho)
    assert np.allclose(prim, p_ref, rtol=1e-5)
    # penalty + gradient
    b = blocks[0]
    grads = [torch.zeros_like(t) for t in b.x]
    for t, gr in zip(b.x, grads):
        t.grad = gr
    pens = b.add_penalty_grads_().cpu().numpy()
    x0 = [t.cpu().numpy() for t in b.x]
    pens_ref, grads_ref = O.admm_penalty(x0, [t.cpu().numpy() for t in b.u], [t.cpu().numpy() for t in b.z], b.rho)
    assert np.allclose(pens, pens_ref, rtol=1e-5, atol=1e-12)
    for g in range(6):
        assert np.abs(grads[g].cpu().numpy() - grads_ref[g]).max() <= 1e-6 * max(1.0, np.abs(grads_ref[g]).max())


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
        plan = A.ConsensusPlan(torch.arange(P, device=DEV), torch.full((P,), 2, dtype=torch.int32, device=DEV))
        x = [lv[k] for k in ("means3D", "dc", "sh", "scales", "rotations", "opacities")]
        blk = A.AdmmBlock(x, plan, alphas=(3e3, 2e3, 5e3, 1e3, 7e3, 9e2))
        gg = torch.Generator(device="cpu").manual_seed(5)
        for t in blk.zmu:                                   # pretend a consensus round happened
            t.add_(0.05 * torch.randn(t.shape, generator=gg).to(DEV))
        blk.fuse_into_rasterizer(fused, all_groups=all_groups)
        try:
            views = dict(lv)
            views["dc"] = lv["dc"].view(P, 1, 3)
            views["sh"] = lv["sh"].view(P, 15, 3)
            img, _, _, _ = scenes.render(dgr, sc, views)
            (img * w).sum().backward()
            pens = blk.add_penalty_grads_()
        finally:
            blk.fuse_into_rasterizer(False)
        torch.cuda.synchronize()
        return {k: v.grad.clone() for k, v in lv.items()}, pens.clone()

    g0, p0 = run(False)
    for all_groups in (False, True):   # True: scaling / quaternion / opacity fused as well (no penalty kernel at all)
        g1, p1 = run(True, all_groups)
        assert torch.allclose(p0, p1, rtol=1e-6, atol=1e-15)
        for k in g0:
            assert U.rel_to_max(g1[k], g0[k]) <= 2e-6, (k, all_groups)


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
