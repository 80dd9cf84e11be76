"""CPU (no GPU): pins the oracle (oracle/dgr_oracle.c) against golden vectors produced by the
UNMODIFIED reference CUDA build on a B200 (tools/make_golden.py -> tests/golden/*.npz).

Bars: per-Gaussian preprocess outputs and everything integer bit-exact; threshold-dependent
quantities that go through expf/logf (libm here, libdevice there) may flip in a handful of places;
images <= 1e-5; gradients <= 1e-4 relative to the tensor's max."""
import os

import numpy as np
import pytest

from oracle import oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RASTER = ["raster_tiny_s0", "raster_tiny_s1_aa_bg"]


def load(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    s = {k[3:]: g[k] for k in g.files if k.startswith("in_")}
    s.update(W=int(g["W"]), H=int(g["H"]), tanfovx=float(g["tanfovx"]), tanfovy=float(g["tanfovy"]),
             sh_degree=int(g["sh_degree"]))
    return g, s


def bits(a):
    return np.ascontiguousarray(a).view(np.int32)


def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.fixture(scope="module", params=RASTER)
def gold(request):
    g, s = load(request.param)
    pre = O.preprocess(s, antialiasing=bool(g["aa"]))
    bins = O.binning(s, pre)
    img = O.render_fwd(s, pre, bins, bg=g["bg"])
    return g, s, pre, bins, img


def test_preprocess_bit_exact(gold):
    g, s, pre, _, _ = gold
    vis = g["radii"] > 0
    assert np.array_equal(pre["radii"], g["radii"])
    for k in ("cov3D", "depths", "means2D", "conic_opacity", "rgb"):
        assert np.array_equal(bits(pre[k][vis]), bits(g[k][vis])), k
    assert np.array_equal(pre["clamped"][vis], g["clamped"][vis])
    assert np.array_equal(pre["tiles_touched"].astype(np.int64), g["tiles_touched"].astype(np.int64))


def test_keys_sort_ranges(gold):
    g, s, pre, bins, _ = gold
    # the cull threshold goes through logf: allow <= 2 flipped tile decisions, exact otherwise
    if len(bins["keys"]) == len(g["keys"]):
        assert np.array_equal(bins["keys"].astype(np.int64), g["keys"])
        assert np.array_equal(bins["point_list"].astype(np.int32), g["point_list"])
        assert np.array_equal(bins["ranges"].astype(np.int32), g["ranges"])
    else:
        assert abs(len(bins["keys"]) - len(g["keys"])) <= 2
    assert int(pre["tiles_touched"].sum()) == int(g["num_rendered_with_pads"])


def test_render_forward(gold):
    g, s, pre, bins, img = gold
    nflip = int((img["n_contrib"].astype(np.int32) != g["n_contrib"]).sum())
    assert nflip <= 4, nflip            # expf ulp differences at the 1/255 and 1e-4 thresholds
    assert np.abs(img["color"] - g["color"]).max() <= 1e-5
    assert np.abs(img["invdepth"] - g["invdepth"]).max() <= 1e-5
    same = img["n_contrib"].astype(np.int32) == g["n_contrib"]
    assert np.abs(img["final_T"][same] - g["final_T"][same]).max() <= 1e-6


@pytest.mark.parametrize("mode", ["intended", "zero_stale"])
def test_backward(gold, mode):
    g, s, pre, bins, img = gold
    got = O.backward(s, pre, bins, img, g["dL_dcolor"], g["dL_dinvdepth"], bg=g["bg"],
                     antialiasing=bool(g["aa"]), ref_quirk=(mode == "zero_stale"))
    for k in ("means3D", "means2D", "dc", "sh", "opacities", "scales", "rotations", "colors", "cov3D"):
        want = g[f"grad_{mode}_{k}"]
        assert rel(got[k].reshape(want.shape), want) <= 1e-4, (k, rel(got[k].reshape(want.shape), want))


def test_visibility_helpers(gold):
    g, s, _, _, _ = gold
    assert np.array_equal(O.mark_visible(s["means3D"], s["viewmatrix"]), g["mark_visible"])
    assert np.array_equal(O.visible_filter(s), g["visible_filter"])


def test_ssim_golden():
    g = np.load(os.path.join(GOLD, "ssim.npz"))
    m, d1, d2, d3 = O.ssim_fwd(g["img1"], g["img2"])
    assert np.abs(m - g["ssim_map"]).max() <= 2e-6
    for a, k in ((d1, "dm_dmu1"), (d2, "dm_dsigma1_sq"), (d3, "dm_dsigma12")):
        assert rel(a, g[k]) <= 1e-5, k
    gi = O.ssim_bwd(g["img1"], g["img2"], g["dL_dmap"], g["dm_dmu1"], g["dm_dsigma1_sq"], g["dm_dsigma12"])
    assert rel(gi, g["dL_dimg1"]) <= 1e-5
    assert abs(float(m.mean()) - float(g["mean_same"])) <= 1e-6
    assert abs(float(m[:, :, 5:-5, 5:-5].mean()) - float(g["mean_valid"])) <= 1e-6


def test_knn_golden():
    g = np.load(os.path.join(GOLD, "knn.npz"))
    d = O.knn_dist2(g["points"])
    assert rel(d, g["dist2"]) <= 1e-6
    assert np.abs(d - g["dist2"]).max() <= 4 * np.spacing(np.float32(g["dist2"].max()))


def test_adam_golden():
    g = np.load(os.path.join(GOLD, "adam.npz"))
    p, m, v = g["param"].copy(), g["exp_avg"].copy(), g["exp_avg_sq"].copy()
    O.sparse_adam(p, g["grad"], m, v, g["visible"], float(g["lr"]), float(g["b1"]), float(g["b2"]), float(g["eps"]))
    assert np.array_equal(bits(m), bits(g["exp_avg_out"]))
    assert np.array_equal(bits(v), bits(g["exp_avg_sq_out"]))
    assert np.array_equal(bits(p), bits(g["param_out"]))
    untouched = ~g["visible"]
    assert np.array_equal(p[untouched], g["param"][untouched])


def test_block_split_oracle_matches_hand_example():
    """select_gaussians_in_each_block (master:124-172) on a case small enough to check by hand."""
    from oracle import oracle as O
    xyz = np.array([[0.5, 0.5, 0], [1.5, 0.5, 0], [2.5, 0.5, 0], [9.0, 9.0, 0], [1.0, 1.0, 0]], np.float32)
    boxes = np.array([[[0, 0, -1], [1.6, 1, 1]], [[1.4, 0, -1], [3, 1, 1]]], np.float64)   # overlap 1.4..1.6
    valid, vis, gidx = O.select_gaussians_in_each_block(xyz, boxes)
    assert valid.tolist() == [0, 1, 2, 4]              # (9, 9) lies in no box
    assert vis.tolist() == [1, 2, 1, 1]                # x = 1.5 is shared; the closed box includes (1, 1)
    assert [g.tolist() for g in gidx] == [[0, 1, 3], [1, 2]]
    # world -> OBB transform: a shift by (-1, 0) moves everything one box to the left
    M = np.array([[1, 0, -1.0], [0, 1, 0], [0, 0, 1]], np.float64)
    valid, vis, gidx = O.select_gaussians_in_each_block(xyz, boxes, M)
    assert valid.tolist() == [1, 2, 4] and [g.tolist() for g in gidx] == [[0, 1, 2], [1]]


def test_l1_ssim_loss_golden():
    g = np.load(os.path.join(GOLD, "loss_l1_ssim.npz"))
    loss, grad = O.l1_ssim_loss(g["img"], g["gt"], float(g["lam"]))
    assert abs(loss - float(g["loss"])) <= 1e-6
    assert rel(grad, g["grad"]) <= 1e-5


def test_activations_golden():
    g = np.load(os.path.join(GOLD, "activations.npz"))
    o, s, q = O.activations(g["raw_opacity"], g["raw_scaling"], g["raw_quaternion"])
    # numpy's float32 exp may differ from CUDA's expf in the last bit; the normalisation is exact arithmetic
    assert np.abs(o - g["opacity"]).max() <= 2e-7 and rel(s, g["scaling"]) <= 2e-7
    assert np.array_equal(q, g["quaternion"])


def test_admm_oracle_matches_reference_methods_golden():
    """oracle.admm_* (numpy restatement of master:326-456,538-555, slave:99-121,161-202 and
    gaussian_splat_model.py:316-340) against tests/golden/admm.npz, produced by executing those reference
    methods themselves (tools/make_golden_admm.py): z and u bit-exact, residuals / penalties to float32
    reduction accuracy, the adapted rho identical."""
    G = dict(np.load(os.path.join(GOLD, "admm.npz")))
    K, n_global = int(G["K"]), int(G["n_global"])
    names = ("xyz", "fdc", "fr", "s", "q", "o")
    gidx = [G[f"global_indices_{k}"] for k in range(K)]
    flat = lambda a: np.ascontiguousarray(a).reshape(a.shape[0], -1)
    rho = list(G["rho_init"])
    assert np.array_equal(np.array(rho), np.array([a / n_global for a in (1e5, 1e4, 1e5, 1e4, 1e5, 1e4)]))
    z_prev = [flat(G[f"z0_{n}"]) for n in names]
    u = [[np.zeros_like(flat(G[f"r0_x_{n}_{k}"])) for n in names] for k in range(K)]
    z = z_prev
    for rnd in range(3):
        xs = [[flat(G[f"r{rnd}_x_{n}_{k}"]) for n in names] for k in range(K)]
        if rnd > 0:
            for k in range(K):
                pens, grads = O.admm_penalty(xs[k], u[k], [zz[gidx[k]] for zz in z], rho)
                assert np.allclose(pens, G[f"r{rnd}_penalty_{k}"], rtol=1e-5, atol=1e-12)
                for gi, n in enumerate(names):
                    assert rel(grads[gi], flat(G[f"r{rnd}_pengrad_{n}_{k}"])) <= 2e-6
        z_prev = z
        z, count = O.admm_consensus(xs, gidx, n_global)
        assert np.array_equal(count.astype(np.int64), G["visibility_count"])
        for gi, n in enumerate(names):
            assert np.array_equal(bits(z[gi]), bits(flat(G[f"r{rnd}_z_{n}"]))), (rnd, n)
        for k in range(K):
            u[k] = O.admm_dual_update(xs[k], u[k], [zz[gidx[k]] for zz in z], 1.5)
            for gi, n in enumerate(names):
                assert np.array_equal(bits(u[k][gi]), bits(flat(G[f"r{rnd}_u_{n}_{k}"]))), (rnd, k, n)
        primal, dual = O.admm_residuals(xs, gidx, z, z_prev, rho)
        assert np.allclose(primal, G[f"r{rnd}_primal"], rtol=2e-6, atol=1e-14)
        assert np.allclose(dual, G[f"r{rnd}_dual"], rtol=2e-6, atol=1e-14)
        rho = O.admm_adapt(rho, primal, dual)
        assert np.array_equal(np.array(rho), G[f"r{rnd}_rho_after"])


def test_densify_oracle_matches_reference_methods_golden():
    """oracle.densify_and_prune / densification_stats / reset_opacity / v_imp_score against tests/golden/densify.npz
    (tools/make_golden_densify.py: the reference's own GaussianSplatModel methods + torch.optim.Adam surgery)."""
    G = dict(np.load(os.path.join(GOLD, "densify.npz")))
    P = int(G["P"])
    names = O.DENSIFY_PARAMS
    accum, denom, mr = np.zeros((P, 1), np.float32), np.zeros((P, 1), np.float32), np.zeros(P, np.float32)
    for it in range(3):
        O.densification_stats(accum, denom, mr, G[f"stat{it}_grad2d"], G[f"stat{it}_radii"])
    # torch.norm's reduction may round the two squares differently (vectorised / fused): 1 ulp
    assert np.abs(accum - G["stats_accum"]).max() <= 3e-7 * np.abs(G["stats_accum"]).max()
    assert np.array_equal(denom, G["stats_denom"])
    accum = G["stats_accum"].copy()
    assert np.array_equal(mr, G["stats_max_radii2D"])
    for case, mss, zmin in (("a", float(G["max_screen_size"]), None), ("b", None, float(G["bbox_zmin"]))):
        p = {n: G[f"in_{n}"] for n in names}
        m = {n: G[f"in_m_{n}"] for n in names}
        v = {n: G[f"in_v_{n}"] for n in names}
        p2, m2, v2 = O.densify_and_prune(p, m, v, accum, denom, G["samples"], float(G["max_grad"]), float(G["min_opacity"]),
                                         float(G["extent"]), mss, bbox_zmin=zmin)
        for n in names:
            ref = G[f"dp_{case}_{n}"]
            assert p2[n].shape == ref.shape, (case, n, p2[n].shape, ref.shape)
            if n in ("xyz", "scaling"):      # split replicas: matmul / log-exp rounding
                assert np.abs(p2[n] - ref).max() <= 2e-6 * max(1.0, np.abs(ref).max()), (case, n)
            else:
                assert np.array_equal(bits(p2[n]), bits(ref)), (case, n)
            assert np.array_equal(bits(m2[n]), bits(G[f"dp_{case}_m_{n}"])), (case, n)
            assert np.array_equal(bits(v2[n]), bits(G[f"dp_{case}_v_{n}"])), (case, n)
        assert not G[f"dp_{case}_accum"].any() and not G[f"dp_{case}_max_radii2D"].any()
    ro = O.reset_opacity(G["dp_a_opacity"])
    assert np.abs(ro - G["reset_opacity"]).max() <= 2e-6 * np.abs(G["reset_opacity"]).max()
    assert not G["reset_m_opacity"].any() and not G["reset_v_opacity"].any()
    assert np.array_equal(G["reset_m_xyz"], G["dp_a_m_xyz"])            # only the opacity group is touched
    vl = O.v_imp_score(G["reset_scaling"], G["imp_list"], 0.1)
    assert np.abs(vl - G["v_list"]).max() <= 1e-5 * np.abs(G["v_list"]).max()
    keep = O.prune_by_score(G["v_list"], 0.6 * 0.5)
    assert np.array_equal(G["reset_xyz"][keep], G["lg_xyz"]) and np.array_equal(G["reset_m_f_rest"][keep], G["lg_m_f_rest"])


@pytest.mark.parametrize("name", ["count_tiny", "count_racefree"])
def test_count_render_oracle_vs_old_rasterizer_golden(name):
    """oracle render_fwd(count=True) (restating renderCUDA_count, old forward.cu:400-513) against the golden of the
    reference's older rasterizer (tools/make_golden_count.py)."""
    path = os.path.join(GOLD, name + ".npz")
    if not os.path.exists(path):
        pytest.skip("golden not generated yet")
    g, s = load(name)
    pre = O.preprocess(s)
    bins = O.binning(s, pre)
    o = O.render_fwd(s, pre, bins, count=True)
    assert np.array_equal(pre["radii"], g["radii"])
    assert np.abs(o["color"] - g["color"]).max() <= 1e-4
    # libm vs libdevice expf: a handful of alpha-threshold decisions may flip
    diff = np.abs(o["gaussians_count"].astype(np.int64) - g["gaussians_count"])
    assert int(diff.sum()) <= max(8, int(g["gaussians_count"].sum()) // 20000)
    same = diff == 0
    # the reference adds `opacity` once per blended pixel in float32 (count roundings); count * opacity has one
    assert np.allclose(o["important_score"][same], g["important_score"][same], rtol=1e-5, atol=1e-7)
