"""numpy front-end of the CPU oracle (oracle/dgr_oracle.c) plus numpy restatements of the ADMM
consensus arithmetic.  TEST INFRASTRUCTURE ONLY — never imported by the product path.

Parity status: the reference has no tests or golden vectors of its own; this oracle
is pinned against tests/golden/*.npz, which were produced by running the unmodified reference
CUDA build (oracle/_ref) on a B200 with tools/make_golden.py."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")
_lib = None


def build(force=False):
    src = os.path.join(HERE, "dgr_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return LIB


TRUTH_LIB = os.path.join(HERE, "libtruth.so")


def build_truth(force=False):
    """FP64 truth of the rasterizer backward (oracle/truth_bwd.cu, CUDA, test infrastructure): cross-compiled
    for sm_100a where nvcc exists; the GPU box uses the prebuilt file."""
    src = os.path.join(HERE, "truth_bwd.cu")
    if force or not os.path.exists(TRUTH_LIB) or os.path.getmtime(TRUTH_LIB) < os.path.getmtime(src):
        nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
        if not os.path.exists(nvcc):
            return None
        subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-shared",
                               "-Xcompiler", "-fPIC", "-o", TRUTH_LIB, src, "-lcudart"])
    return TRUTH_LIB


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        _lib.orc_duplicate_keys.restype = C.c_int64
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def num_threads():
    return int(os.environ.get("OMP_NUM_THREADS", os.cpu_count() or 1))


def preprocess(s, antialiasing=False):
    """forward.cu:156-276.  s: scene dict of numpy arrays (see dogs_b200.scenes.make_scene)."""
    P = s["means3D"].shape[0]
    M = s["sh"].shape[1] if s.get("sh") is not None else 0
    o = dict(radii=np.zeros(P, np.int32), means2D=np.zeros((P, 2), np.float32),
             depths=np.zeros(P, np.float32), cov3D=np.zeros((P, 6), np.float32),
             rgb=np.zeros((P, 3), np.float32), conic_opacity=np.zeros((P, 4), np.float32),
             clamped=np.zeros((P, 3), np.uint8), tiles_touched=np.zeros(P, np.uint32))
    lib().orc_preprocess(
        C.c_int(P), C.c_int(int(s["sh_degree"])), C.c_int(M), _p(_f32(s["means3D"])),
        _p(_f32(s.get("scales"))), C.c_float(s.get("scale_modifier", 1.0)),
        _p(_f32(s.get("rotations"))), _p(_f32(s["opacities"])), _p(_f32(s.get("dc"))),
        _p(_f32(s.get("sh"))), _p(_f32(s.get("cov3D_precomp"))), _p(_f32(s.get("colors_precomp"))),
        _p(_f32(s["viewmatrix"])), _p(_f32(s["projmatrix"])), _p(_f32(s["campos"])),
        C.c_int(int(s["W"])), C.c_int(int(s["H"])), C.c_float(s["tanfovx"]), C.c_float(s["tanfovy"]),
        C.c_int(int(antialiasing)), _p(o["radii"]), _p(o["means2D"]), _p(o["depths"]), _p(o["cov3D"]),
        _p(o["rgb"]), _p(o["conic_opacity"]), _p(o["clamped"]), _p(o["tiles_touched"]))
    return o


def binning(s, pre):
    """rasterizer_impl.cu:120-220: keys (pads dropped), stable sort, tile ranges."""
    P = pre["radii"].shape[0]
    W, H = int(s["W"]), int(s["H"])
    T = ((W + 15) // 16) * ((H + 15) // 16)
    cap = int(pre["tiles_touched"].sum())
    keys = np.zeros(max(cap, 1), np.uint64)
    vals = np.zeros(max(cap, 1), np.uint32)
    counts = np.zeros(P, np.uint32)
    n = lib().orc_duplicate_keys(C.c_int(P), _p(pre["means2D"]), _p(pre["conic_opacity"]),
                                 _p(pre["depths"]), _p(pre["radii"]), C.c_int(W), C.c_int(H),
                                 _p(keys), _p(vals), _p(counts))
    keys, vals = keys[:n].copy(), vals[:n].copy()
    unsorted_keys = keys.copy()
    lib().orc_sort_pairs(C.c_int64(n), _p(keys), _p(vals))
    ranges = np.zeros((T, 2), np.uint32)
    lib().orc_tile_ranges(C.c_int64(n), _p(keys), C.c_int(T), _p(ranges))
    return dict(keys=keys, point_list=vals, ranges=ranges, counts=counts, keys_unsorted=unsorted_keys)


def render_fwd(s, pre, bins, bg=None, tile_sel=None, count=False):
    """forward.cu:349-501."""
    W, H = int(s["W"]), int(s["H"])
    bg = np.zeros(3, np.float32) if bg is None else _f32(bg)
    o = dict(color=np.zeros((3, H, W), np.float32), invdepth=np.zeros((1, H, W), np.float32),
             final_T=np.zeros(H * W, np.float32), n_contrib=np.zeros(H * W, np.uint32))
    cnt = np.zeros(pre["radii"].shape[0], np.int32) if count else None
    lib().orc_render_fwd(C.c_int(W), C.c_int(H), _p(bins["ranges"]), _p(bins["point_list"]),
                         _p(pre["means2D"]), _p(pre["rgb"]), _p(pre["conic_opacity"]),
                         _p(pre["depths"]), _p(bg), _p(o["color"]), _p(o["invdepth"]),
                         _p(o["final_T"]), _p(o["n_contrib"]), _p(tile_sel), _p(cnt))
    if count:
        # old_diff-gaussian-rasterization forward.cu:485-486: count++ and score += conic_opacity.w per blended pixel
        o["gaussians_count"] = cnt
        o["important_score"] = (cnt.astype(np.float32) * pre["conic_opacity"][:, 3]).astype(np.float32)
    return o


def forward(s, bg=None, antialiasing=False):
    pre = preprocess(s, antialiasing)
    bins = binning(s, pre)
    img = render_fwd(s, pre, bins, bg)
    return pre, bins, img


def backward(s, pre, bins, img, dL_dcolor, dL_dinvdepth=None, bg=None, antialiasing=False,
             depth_threshold=0.0, ref_quirk=False, tile_sel=None):
    """backward.cu (render + cov2D + preprocess) -> dict of gradients shaped like the reference's."""
    P = pre["radii"].shape[0]
    W, H = int(s["W"]), int(s["H"])
    M = s["sh"].shape[1]
    bg = np.zeros(3, np.float32) if bg is None else _f32(bg)
    acc = np.zeros((P, 10), np.float64)
    dLc = _f32(dL_dcolor)
    dLd = _f32(dL_dinvdepth)
    lib().orc_render_bwd(C.c_int(W), C.c_int(H), _p(bins["ranges"]), _p(bins["point_list"]), _p(bg),
                         _p(pre["means2D"]), _p(pre["conic_opacity"]), _p(pre["rgb"]),
                         _p(pre["depths"]), _p(img["final_T"]), _p(img["n_contrib"]),
                         _p(img["color"]), _p(img["invdepth"]), _p(dLc), _p(dLd), _p(acc),
                         C.c_int(int(ref_quirk)), _p(tile_sel))
    g = dict(means3D=np.zeros((P, 3), np.float32), means2D=np.zeros((P, 3), np.float32),
             dc=np.zeros((P, 1, 3), np.float32), sh=np.zeros((P, M, 3), np.float32),
             colors=np.zeros((P, 3), np.float32), opacities=np.zeros((P, 1), np.float32),
             scales=np.zeros((P, 3), np.float32), rotations=np.zeros((P, 4), np.float32),
             cov3D=np.zeros((P, 6), np.float32))
    lib().orc_preprocess_bwd(
        C.c_int(P), C.c_int(int(s["sh_degree"])), C.c_int(M), _p(_f32(s["means3D"])), _p(pre["radii"]),
        _p(_f32(s["dc"])), _p(_f32(s["sh"])), _p(pre["clamped"]), _p(_f32(s["opacities"])),
        _p(_f32(s["scales"])), _p(_f32(s["rotations"])), C.c_float(s.get("scale_modifier", 1.0)), None,
        _p(_f32(s["viewmatrix"])), _p(_f32(s["projmatrix"])), C.c_int(W), C.c_int(H),
        C.c_float(s["tanfovx"]), C.c_float(s["tanfovy"]), _p(_f32(s["campos"])), _p(acc),
        C.c_int(int(antialiasing)), C.c_float(depth_threshold), _p(g["means3D"]), _p(g["means2D"]),
        _p(g["dc"]), _p(g["sh"]), _p(g["colors"]), _p(g["opacities"]), _p(g["scales"]),
        _p(g["rotations"]), _p(g["cov3D"]))
    return g


def visible_filter(s):
    P = s["means3D"].shape[0]
    radii = np.zeros(P, np.int32)
    lib().orc_visible_filter(C.c_int(P), _p(_f32(s["means3D"])), _p(_f32(s["scales"])),
                             C.c_float(s.get("scale_modifier", 1.0)), _p(_f32(s["rotations"])), None,
                             _p(_f32(s["viewmatrix"])), _p(_f32(s["projmatrix"])), C.c_int(int(s["W"])),
                             C.c_int(int(s["H"])), C.c_float(s["tanfovx"]), C.c_float(s["tanfovy"]),
                             _p(radii))
    return radii


def mark_visible(means, view):
    P = means.shape[0]
    out = np.zeros(P, np.uint8)
    lib().orc_mark_visible(C.c_int(P), _p(_f32(means)), _p(_f32(view)), _p(out))
    return out.astype(bool)


def ssim_fwd(img1, img2, C1=0.01 ** 2, C2=0.03 ** 2, train=True):
    a, b = _f32(img1), _f32(img2)
    B, CH, H, W = a.shape
    m = np.zeros_like(a)
    d = [np.zeros_like(a) for _ in range(3)] if train else [None] * 3
    lib().orc_ssim_fwd(C.c_int(B * CH), C.c_int(H), C.c_int(W), C.c_float(C1), C.c_float(C2), _p(a),
                       _p(b), _p(m), _p(d[0]), _p(d[1]), _p(d[2]))
    return m, d[0], d[1], d[2]


def ssim_bwd(img1, img2, dL_dmap, d1, d2, d3):
    a, b, g = _f32(img1), _f32(img2), _f32(dL_dmap)
    B, CH, H, W = a.shape
    out = np.zeros_like(a)
    lib().orc_ssim_bwd(C.c_int(B * CH), C.c_int(H), C.c_int(W), _p(a), _p(b), _p(g), _p(_f32(d1)),
                       _p(_f32(d2)), _p(_f32(d3)), _p(out))
    return out


def knn_dist2(pts):
    p = _f32(pts)
    out = np.zeros(p.shape[0], np.float32)
    lib().orc_knn_dist2(C.c_int(p.shape[0]), _p(p), _p(out))
    return out


def sparse_adam(param, grad, m, v, visible, lr, b1, b2, eps):
    """In place on float32 arrays (adam.cu:9-38)."""
    N = visible.shape[0]
    M = param.size // N
    vis = np.ascontiguousarray(visible, dtype=np.uint8)
    lib().orc_sparse_adam(_p(param), _p(_f32(grad)), _p(m), _p(v), _p(vis), C.c_float(lr),
                          C.c_float(b1), C.c_float(b2), C.c_float(eps), C.c_uint32(N), C.c_uint32(M))


# ---- ADMM (numpy restatement of the reference's torch arithmetic; pinned by tests/golden/admm.npz, which
# tools/make_golden_admm.py produced by executing the reference's own master / slave methods) ----------
GROUPS = ("xyz", "fdc", "fr", "s", "q", "o")
GROUP_SIZES = (3, 3, 45, 3, 4, 1)


def admm_consensus(block_x, global_indices, n_global):
    """master_gaussian_trainer.py:538-555 + gaussian_splat_model.py:316-340:
    z = (sum_k scatter(x_k)) / count.  block_x: list over blocks of list of 6 arrays [n_k, size]."""
    z = [np.zeros((n_global, sz), np.float32) for sz in GROUP_SIZES]
    count = np.zeros(n_global, np.float32)
    for xk, idx in zip(block_x, global_indices):
        for g in range(6):
            z[g][idx] += xk[g].reshape(len(idx), -1)       # plus_gaussians (indices unique per block)
        count[idx] += 1
    for g in range(6):
        z[g] = (z[g] / count[:, None]).astype(np.float32)   # average_gaussians
    return z, count


def admm_dual_update(x, u, z_local, relax=1.5):
    """slave_gaussian_trainer.py:99-121: u += (1 + alpha) * (x - z)."""
    return [(ug + np.float32(relax) * (xg - zg)).astype(np.float32) for xg, ug, zg in zip(x, u, z_local)]


def admm_penalty(x, u, z_local, rho):
    """slave_gaussian_trainer.py:161-202: sum_g 0.5*rho_g*mse(x_g+u_g, z_g) and its gradient."""
    pens, grads = [], []
    for g in range(6):
        r = (x[g] + u[g]) - z_local[g]
        pens.append(0.5 * rho[g] * float(np.mean(r.astype(np.float64) ** 2)))
        grads.append((rho[g] / r.size * r).astype(np.float32))
    return pens, grads


def admm_residuals(block_x, global_indices, z, z_prev, rho):
    """master:396-456: primal_g = sum_k mse(z[idx_k], x_k); dual_g = rho_g * mse(z_prev, z)."""
    primal = [0.0] * 6
    for xk, idx in zip(block_x, global_indices):
        for g in range(6):
            d = z[g][idx].astype(np.float64) - xk[g].reshape(len(idx), -1)
            primal[g] += float(np.mean(d ** 2))
    dual = [rho[g] * float(np.mean((z_prev[g].astype(np.float64) - z[g]) ** 2)) for g in range(6)]
    return primal, dual


def admm_adapt(rho, primal, dual, mu=10.0, tau_inc=2.0, tau_dec=2.0):
    """master:337-377."""
    out = list(rho)
    for g in range(6):
        if primal[g] > mu * dual[g]:
            out[g] *= tau_inc
        elif dual[g] > mu * primal[g]:
            out[g] /= tau_dec
    return out


def points_in_bbox2d(points, bbox, transform_world_to_obb=None):
    """conerf/datasets/utils.py:186-205.  points [N,2] (float32, as model.get_xyz.cpu().numpy()[:, :2]),
    bbox [2,3] min/max corners; trimesh.transform_points on 2-D points = (M . [x, y, 1]^T)[:2] in
    float64, no perspective divide."""
    A, B = bbox[0, :], bbox[1, :]
    pts = points
    if transform_world_to_obb is not None:
        M = np.asarray(transform_world_to_obb, np.float64)
        stack = np.column_stack((points, np.ones(len(points))))
        pts = np.dot(M, stack.T).T[:, :2]
    return np.argwhere((A[0] <= pts[:, 0]) & (pts[:, 0] <= B[0]) & (A[1] <= pts[:, 1]) & (pts[:, 1] <= B[1])).reshape(-1)


def select_gaussians_in_each_block(xyz, bboxes, transform_world_to_obb=None):
    """master_gaussian_trainer.py:124-172: drop Gaussians that lie in no box, regroup the rest per box.
    Returns (valid_gaussian_indices, visibility_count, global_indices)."""
    pts = np.asarray(xyz, np.float32)[:, :2]
    first = [points_in_bbox2d(pts, np.asarray(b, np.float64).reshape(2, 3), transform_world_to_obb) for b in bboxes]
    visibility = np.bincount(np.concatenate(first), minlength=len(pts))
    valid = np.argwhere(visibility).reshape(-1)
    pts = pts[valid]
    global_indices = [points_in_bbox2d(pts, np.asarray(b, np.float64).reshape(2, 3), transform_world_to_obb)
                      for b in bboxes]
    visibility = np.bincount(np.concatenate(global_indices), minlength=len(pts))
    assert (visibility != 0).all()
    return valid, visibility, global_indices


def l1_ssim_loss(img, gt, lam=0.2):
    """conerf/trainers/gaussian_trainer.py:403-405: (1-l)*l1_loss(img, gt) + l*(1 - fused_ssim(img, gt)) and its
    gradient w.r.t. img, composed from the SSIM restatement above.  img, gt: [C,H,W] float32."""
    a, b = _f32(img)[None], _f32(gt)[None]
    m, d1, d2, d3 = ssim_fwd(a, b)
    n = a.size
    loss = (1.0 - lam) * float(np.abs(a.astype(np.float64) - b).mean()) + lam * (1.0 - float(m.astype(np.float64).mean()))
    dmap = np.full_like(a, np.float32(-lam) / np.float32(n))
    grad = ssim_bwd(a, b, dmap, d1, d2, d3) + (np.float32(1.0 - lam) / np.float32(n)) * np.sign(a - b)
    return loss, grad[0].astype(np.float32)


def activations(raw_opacity, raw_scaling, raw_quaternion):
    """gaussian_splat_model.py:148-153: sigmoid, exp, F.normalize(dim=1, eps=1e-12) in float32."""
    o = (np.float32(1) / (np.float32(1) + np.exp(-_f32(raw_opacity)))).astype(np.float32)
    s = np.exp(_f32(raw_scaling)).astype(np.float32)
    q = _f32(raw_quaternion)
    sq = q * q
    nrm = np.sqrt((sq[:, 0] + sq[:, 2]) + (sq[:, 1] + sq[:, 3]))       # torch's row-reduction order
    return o, s, (q / np.maximum(nrm, np.float32(1e-12))[:, None]).astype(np.float32)



# ---- densification / pruning / opacity reset (numpy restatement of GaussianSplatModel's methods; pinned by
# tests/golden/densify.npz, which tools/make_golden_densify.py produced by executing those methods) ----------
DENSIFY_PARAMS = ("xyz", "f_dc", "f_rest", "opacity", "scaling", "quaternion")


def densification_stats(accum, denom, max_radii2D, grad2d, radii):
    """gaussian_trainer.py:433-439 + add_densification_stats (gaussian_splat_model.py:529-541), in place."""
    vis = radii > 0
    max_radii2D[vis] = np.maximum(max_radii2D[vis], radii[vis].astype(np.float32))
    g = _f32(grad2d)
    accum[vis, 0] += np.sqrt(g[vis, 0] * g[vis, 0] + g[vis, 1] * g[vis, 1]).astype(np.float32)
    denom[vis, 0] += np.float32(1)


def _quat_to_rot(q):
    """conerf/model/gaussian_fields/utils.py:27-70."""
    n = np.sqrt(q[:, 0] * q[:, 0] + q[:, 1] * q[:, 1] + q[:, 2] * q[:, 2] + q[:, 3] * q[:, 3])
    q = q / n[:, None]
    r, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = np.zeros((q.shape[0], 3, 3), np.float32)
    R[:, 0, 0] = 1 - 2 * (y * y + z * z); R[:, 0, 1] = 2 * (x * y - r * z); R[:, 0, 2] = 2 * (x * z + r * y)
    R[:, 1, 0] = 2 * (x * y + r * z); R[:, 1, 1] = 1 - 2 * (x * x + z * z); R[:, 1, 2] = 2 * (y * z - r * x)
    R[:, 2, 0] = 2 * (x * z - r * y); R[:, 2, 1] = 2 * (y * z + r * x); R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def densify_and_prune(p, m, v, accum, denom, samples, max_grad, min_opacity, extent, max_screen_size,
                      percent_dense=0.01, bbox_zmin=None, num_replica=2):
    """densify_and_prune (:499-527) with densify_and_clone (:433-453), densify_and_split (:455-497),
    densification_postfix (:369-404), prune_points (:406-416) and the optimizer surgery (:34-108) on dicts of
    numpy arrays: p parameters, m / v Adam moments.  `samples` = the torch.normal draw of :472-474.
    Returns (p, m, v) of the new model; its statistics are all zero (postfix)."""
    f = np.float32
    sig = lambda o: (f(1) / (f(1) + np.exp(-o))).astype(f)
    with np.errstate(invalid="ignore", divide="ignore"):
        grads = (accum / denom).astype(f)
    grads[np.isnan(grads)] = 0
    cat = lambda d, extra: {k: np.concatenate([d[k], extra[k]], 0) for k in DENSIFY_PARAMS}
    zeros_like = lambda d: {k: np.zeros_like(d[k]) for k in DENSIFY_PARAMS}
    take = lambda d, idx: {k: d[k][idx] for k in DENSIFY_PARAMS}
    thr_scale = f(percent_dense * extent)
    # clone
    sel = (np.abs(grads[:, 0]) >= f(max_grad)) & (np.exp(p["scaling"]).max(1) <= thr_scale)
    new = take(p, sel)
    p, m, v = cat(p, new), cat(m, zeros_like(new)), cat(v, zeros_like(new))
    # split
    n_init = p["xyz"].shape[0]
    padded = np.zeros(n_init, f)
    padded[:grads.shape[0]] = grads[:, 0]
    sel = (padded >= f(max_grad)) & (np.exp(p["scaling"]).max(1) > thr_scale)
    R = np.tile(_quat_to_rot(p["quaternion"][sel]), (num_replica, 1, 1))
    new = {k: np.tile(p[k][sel], (num_replica,) + (1,) * (p[k].ndim - 1)) for k in DENSIFY_PARAMS}
    new["xyz"] = (np.einsum("nij,nj->ni", R, _f32(samples)) + np.tile(p["xyz"][sel], (num_replica, 1))).astype(f)
    new["scaling"] = np.log(np.tile(np.exp(p["scaling"][sel]), (num_replica, 1)) / f(0.8 * num_replica)).astype(f)
    p, m, v = cat(p, new), cat(m, zeros_like(new)), cat(v, zeros_like(new))
    keep = ~np.concatenate([sel, np.zeros(num_replica * int(sel.sum()), bool)])
    p, m, v = take(p, keep), take(m, keep), take(v, keep)
    # final prune (max_radii2D was zeroed by the postfix, so the screen-size test never fires)
    prune = sig(p["opacity"])[:, 0] < f(min_opacity)
    if bbox_zmin is not None:
        prune |= p["xyz"][:, 2] < f(bbox_zmin)
    if max_screen_size is not None:
        prune |= np.exp(p["scaling"]).max(1) > f(0.1 * extent)
    keep = ~prune
    return take(p, keep), take(m, keep), take(v, keep)


def reset_opacity(opacity, cap=0.01):
    """reset_opacity (:362-367): inverse_sigmoid(min(sigmoid(o), 0.01)); the Adam moments become zero."""
    f = np.float32
    x = np.minimum((f(1) / (f(1) + np.exp(-opacity))).astype(f), f(cap))
    return np.log(x / (f(1) - x)).astype(f)


def v_imp_score(scaling_raw, imp_list, v_pow):
    """prune.py:14-31."""
    volume = np.prod(np.exp(scaling_raw), axis=1)
    kth = np.sort(volume)[::-1][int(len(volume) * 0.9)]
    return (np.power(volume / kth, np.float32(v_pow)) * imp_list).astype(np.float32)


def prune_by_score(score, percent):
    """prune_gaussians_with_opt (:418-423): keep mask."""
    s = np.sort(score)
    return ~(score <= s[int(percent * (len(s) - 1))])
