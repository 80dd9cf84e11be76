"""CPU baseline named by BASELINE.json's north_star: a float32 PyTorch evaluation of the rasterizer math
(preprocess -> per-tile depth sort -> per-tile blend -> photometric loss -> autograd backward) on the host
cores.  TEST / BENCH INFRASTRUCTURE ONLY (bench.py's cpu_baseline leg); never imported by the product path.

It restates, vectorised over Gaussians and over (256 pixels x tile list), what the reference computes in
cuda_rasterizer/forward.cu:24-276 (preprocess, SH colours), rasterizer_impl.cu:120-220 (tile lists; the
rect only, without the per-tile reachability cull) and forward.cu:412-501 (front-to-back blend with the
1/255 and 1e-4 thresholds); the backward is torch autograd.  All P Gaussians are preprocessed (forward and
backward); the blend runs on a random sample of tiles and is scaled by T / sample.
"""
import math
import os
import time

import torch

SH_C0 = 0.28209479177387814
SH_C1 = 0.4886025119029199
SH_C2 = (1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396)
SH_C3 = (-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
         1.445305721320277, -0.5900435899266435)


def preprocess(p, sc):
    """forward.cu:156-276 for all Gaussians.  Returns dict of [P, ...] tensors (autograd-tracked)."""
    W, H = int(sc["W"]), int(sc["H"])
    view, proj = sc["viewmatrix"], sc["projmatrix"]          # column-major: row-vector @ matrix
    mean = p["means3D"]
    ones = torch.ones_like(mean[:, :1])
    hom = torch.cat([mean, ones], 1)
    p_hom = hom @ proj
    p_w = 1.0 / (p_hom[:, 3] + 1e-7)
    ndc = p_hom[:, :2] * p_w[:, None]
    t = hom @ view
    tz = t[:, 2]
    visible = tz > 0.2
    # cov3D = (S R)^T (S R)
    q = p["rotations"]
    r, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
                     2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
                     2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], 1).view(-1, 3, 3)
    M = R * p["scales"][:, None, :]                          # R diag(s)
    Sigma = M @ M.transpose(1, 2)
    fx, fy = W / (2.0 * sc["tanfovx"]), H / (2.0 * sc["tanfovy"])
    limx, limy = 1.3 * sc["tanfovx"], 1.3 * sc["tanfovy"]
    tzc = tz.clamp_min(0.2)
    tx = torch.clamp(t[:, 0] / tzc, -limx, limx) * tzc
    ty = torch.clamp(t[:, 1] / tzc, -limy, limy) * tzc
    zero = torch.zeros_like(tz)
    J = torch.stack([fx / tzc, zero, -(fx * tx) / (tzc * tzc), zero, fy / tzc, -(fy * ty) / (tzc * tzc)], 1).view(-1, 2, 3)
    Wv = view[:3, :3].t()
    T = J @ Wv
    cov = T @ Sigma @ T.transpose(1, 2)
    a, b, c = cov[:, 0, 0] + 0.3, cov[:, 0, 1], cov[:, 1, 1] + 0.3
    det = a * c - b * b
    det_inv = 1.0 / det
    conic = torch.stack([c * det_inv, -b * det_inv, a * det_inv], 1)
    mid = 0.5 * (a + c)
    lam = mid + torch.sqrt(torch.clamp(mid * mid - det, min=0.1))
    radius = torch.ceil(3.0 * torch.sqrt(lam)).detach()
    xy = torch.stack([((ndc[:, 0] + 1.0) * W - 1.0) * 0.5, ((ndc[:, 1] + 1.0) * H - 1.0) * 0.5], 1)
    # SH colours, degree 3 (forward.cu:24-76)
    d = mean - sc["campos"][None, :]
    d = d / d.norm(dim=1, keepdim=True)
    dx, dy, dz = d[:, 0:1], d[:, 1:2], d[:, 2:3]
    sh = p["sh"]
    col = SH_C0 * p["dc"][:, 0] - SH_C1 * dy * sh[:, 0] + SH_C1 * dz * sh[:, 1] - SH_C1 * dx * sh[:, 2]
    xx, yy, zz, xy_, yz, xz = dx * dx, dy * dy, dz * dz, dx * dy, dy * dz, dx * dz
    col = col + SH_C2[0] * xy_ * sh[:, 3] + SH_C2[1] * yz * sh[:, 4] + SH_C2[2] * (2 * zz - xx - yy) * sh[:, 5] + \
        SH_C2[3] * xz * sh[:, 6] + SH_C2[4] * (xx - yy) * sh[:, 7]
    col = col + SH_C3[0] * dy * (3 * xx - yy) * sh[:, 8] + SH_C3[1] * xy_ * dz * sh[:, 9] + \
        SH_C3[2] * dy * (4 * zz - xx - yy) * sh[:, 10] + SH_C3[3] * dz * (2 * zz - 3 * xx - 3 * yy) * sh[:, 11] + \
        SH_C3[4] * dx * (4 * zz - xx - yy) * sh[:, 12] + SH_C3[5] * dz * (xx - yy) * sh[:, 13] + \
        SH_C3[6] * dx * (xx - 3 * yy) * sh[:, 14]
    col = torch.clamp_min(col + 0.5, 0.0)
    return dict(xy=xy, conic=conic, opacity=p["opacities"][:, 0], rgb=col, depth=tz, radius=radius, visible=visible)


def blend_tile(pre, ids, tx, ty, bg=0.0):
    """forward.cu:412-501 for one 16x16 tile, vectorised over [256 pixels, L list entries]."""
    dev = pre["xy"].device
    py, px = torch.meshgrid(torch.arange(16, dtype=torch.float32, device=dev) + 16 * ty,
                            torch.arange(16, dtype=torch.float32, device=dev) + 16 * tx, indexing="ij")
    pix = torch.stack([px.reshape(-1), py.reshape(-1)], 1)                     # [256, 2]
    order = ids[torch.argsort(pre["depth"][ids].detach(), stable=True)]
    xy, con, op, rgb = pre["xy"][order], pre["conic"][order], pre["opacity"][order], pre["rgb"][order]
    dxy = xy[None, :, :] - pix[:, None, :]                                     # [256, L, 2]
    power = -0.5 * (con[None, :, 0] * dxy[..., 0] ** 2 + con[None, :, 2] * dxy[..., 1] ** 2) - con[None, :, 1] * dxy[..., 0] * dxy[..., 1]
    alpha = torch.clamp_max(op[None, :] * torch.exp(power), 0.99)
    alpha = torch.where((power > 0) | (alpha < 1.0 / 255.0), torch.zeros_like(alpha), alpha)
    Tcum = torch.cumprod(1.0 - alpha, dim=1)                                   # T after each entry
    Tbefore = torch.cat([torch.ones_like(Tcum[:, :1]), Tcum[:, :-1]], 1)
    alive = (Tcum >= 1e-4).to(alpha.dtype).cumprod(dim=1)                      # stops at the first T' < 1e-4
    w = alpha * Tbefore * alive
    color = w @ rgb                                                            # [256, 3]
    return color.t().reshape(3, 16, 16)


def _ssim(a, b):
    """conerf/loss/ssim_torch.py:26-52 (11-tap Gaussian window, sigma 1.5, zero padding) on [B,3,h,w]."""
    g = torch.tensor([math.exp(-(i - 5) ** 2 / (2 * 1.5 ** 2)) for i in range(11)], dtype=a.dtype)
    g = g / g.sum()
    w = (g[:, None] @ g[None, :])[None, None].expand(3, 1, 11, 11).contiguous()
    conv = lambda t: torch.nn.functional.conv2d(t, w, padding=5, groups=3)
    mu1, mu2 = conv(a), conv(b)
    s1, s2, s12 = conv(a * a) - mu1 * mu1, conv(b * b) - mu2 * mu2, conv(a * b) - mu1 * mu2
    c1, c2 = 0.01 ** 2, 0.03 ** 2
    return (((2 * mu1 * mu2 + c1) * (2 * s12 + c2)) / ((mu1 * mu1 + mu2 * mu2 + c1) * (s1 + s2 + c2))).mean()


def measure(scene_cpu, n_tiles=384, seed=0, unit="iters/s"):
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    sc = scene_cpu
    W, H = int(sc["W"]), int(sc["H"])
    gx, gy = (W + 15) // 16, (H + 15) // 16
    T = gx * gy
    keys = ("means3D", "dc", "sh", "opacities", "scales", "rotations")
    p = {k: sc[k].clone().requires_grad_(True) for k in keys}
    g = torch.Generator(device="cpu").manual_seed(seed)
    sel = torch.randperm(T, generator=g)[:min(n_tiles, T)].tolist()

    t0 = time.perf_counter()
    pre = preprocess(p, sc)
    t1 = time.perf_counter()
    # tile lists from the rects (auxiliary.h:45-55), for the sampled tiles only
    xy, rad = pre["xy"].detach(), pre["radius"]
    x0 = ((xy[:, 0] - rad) / 16).floor().clamp(0, gx)
    x1 = ((xy[:, 0] + rad + 15) / 16).floor().clamp(0, gx)
    y0 = ((xy[:, 1] - rad) / 16).floor().clamp(0, gy)
    y1 = ((xy[:, 1] + rad + 15) / 16).floor().clamp(0, gy)
    vis = pre["visible"] & (pre["opacity"].detach() >= 1.0 / 255.0)
    lists = []
    for t in sel:
        ty, tx = divmod(t, gx)
        lists.append(torch.nonzero(vis & (x0 <= tx) & (tx < x1) & (y0 <= ty) & (ty < y1)).squeeze(1))
    t2 = time.perf_counter()
    imgs = torch.stack([blend_tile(pre, ids, t % gx, t // gx) for t, ids in zip(sel, lists)])       # [S,3,16,16]
    target = torch.rand(imgs.shape, generator=g)
    loss = 0.8 * (imgs - target).abs().mean() + 0.2 * (1.0 - _ssim(imgs, target))
    t3 = time.perf_counter()
    loss.backward()
    t4 = time.perf_counter()
    # the per-Gaussian share of that backward: autograd of the preprocess graph alone
    for v in p.values():
        v.grad = None
    pre2 = preprocess(p, sc)
    dummy = pre2["conic"].sum() + pre2["xy"].sum() + pre2["rgb"].sum() + pre2["opacity"].sum()
    t5 = time.perf_counter()
    dummy.backward()
    t6 = time.perf_counter()
    pre_fwd, lists_s, tiles_fwd = t1 - t0, t2 - t1, t3 - t2
    pre_bwd = t6 - t5
    tiles_bwd = max((t4 - t3) - pre_bwd, 0.0)
    scale = T / float(len(sel))
    total = pre_fwd + pre_bwd + scale * (lists_s + tiles_fwd + tiles_bwd)
    return {"value": 1.0 / total, "unit": unit, "cores": threads, "kind": "torch",
            "sample": f"float32 PyTorch, {threads} threads: all {sc['means3D'].shape[0]} Gaussians preprocessed forward+backward "
                      f"({pre_fwd:.2f}+{pre_bwd:.2f} s), tile lists + blend + loss + autograd backward on {len(sel)} of {T} tiles "
                      f"({lists_s + tiles_fwd + tiles_bwd:.2f} s) scaled x{scale:.1f}; {t6 - t0:.1f} s of CPU work",
            "est_s_per_iter": total, "mean_list_length": float(sum(len(i) for i in lists)) / max(len(lists), 1)}
