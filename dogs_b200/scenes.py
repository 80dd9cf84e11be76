"""Seeded synthetic Gaussian scenes of the BASELINE.json configurations and a
restatement of the reference's `render()` glue (conerf/render/gaussian_render.py:18-158) that can
drive either rasterizer module (this package's or the reference build under oracle/_ref)."""
import math
from typing import NamedTuple

import torch


class SceneSpec(NamedTuple):
    name: str
    P: int
    W: int
    H: int
    fx: float          # focal length in pixels (fy = fx)
    eye_z: float
    mean_kind: str     # "normal" | "slab"
    mean_args: tuple
    log_scale_mu: float
    log_scale_sigma: float


def _fx_from_fov(fov, pixels):
    return pixels / (2.0 * math.tan(fov / 2.0))


SPECS = {
    # C1 blender-like: 100K Gaussians, 800x800, fov 0.6911 rad
    "C1": SceneSpec("C1", 100_000, 800, 800, _fx_from_fov(0.6911, 800), -4.0, "normal", (0.6,),
                    math.log(0.010), 0.5),
    # C2 mip360-shaped: 1M Gaussians, 1297x840
    "C2": SceneSpec("C2", 1_000_000, 1297, 840, 1040.0, -4.5, "slab", (2.6, 1.7, 0.5),
                    math.log(0.008), 0.7),
    # C3 1080p headline: 3M Gaussians, 1920x1080
    "C3": SceneSpec("C3", 3_000_000, 1920, 1080, 1500.0, -5.0, "slab", (3.0, 1.7, 0.5),
                    math.log(0.006), 0.7),
    # small case for smoke / unit tests
    "tiny": SceneSpec("tiny", 2_000, 160, 112, 140.0, -4.0, "normal", (0.6,), math.log(0.03), 0.5),
}


def focal_length_to_fov(focal_length, pixels):  # conerf/geometry/camera.py:79-80
    return 2 * math.atan(pixels / (2 * focal_length))


def projection_matrix(znear, zfar, fov_x, fov_y):  # conerf/geometry/pose_util.py:428-448
    tan_half_fov_y = math.tan(fov_y / 2)
    tan_half_fov_x = math.tan(fov_x / 2)
    top = tan_half_fov_y * znear
    bottom = -top
    right = tan_half_fov_x * znear
    left = -right
    P = torch.zeros(4, 4)
    P[0, 0] = 2.0 * znear / (right - left)
    P[1, 1] = 2.0 * znear / (top - bottom)
    P[0, 2] = (right + left) / (right - left)
    P[1, 2] = (top + bottom) / (top - bottom)
    P[3, 2] = 1.0
    P[2, 2] = zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    return P


def make_camera(W, H, fx, fy, world_to_camera):
    """Matrices exactly as conerf/geometry/camera.py:117-135 builds them (transposed = column-major)."""
    fov_x = focal_length_to_fov(fx, W)
    fov_y = focal_length_to_fov(fy, H)
    view = world_to_camera.clone().float().transpose(0, 1).contiguous()
    proj = projection_matrix(0.01, 100.0, fov_x, fov_y).transpose(0, 1)
    full = (view @ proj).contiguous()
    campos = view.inverse()[3, :3].contiguous()
    return dict(W=W, H=H, tanfovx=math.tan(fov_x * 0.5), tanfovy=math.tan(fov_y * 0.5),
                viewmatrix=view, projmatrix=full, campos=campos)


def make_scene(cfg="C1", seed=0, device="cpu", P=None, sh_degree=3):
    """Returns dict(means3D, dc, sh, opacities, scales, rotations, camera...) of float32 tensors.
    Activated values (post sigmoid/exp/normalize), as render() hands them to the rasterizer."""
    spec = SPECS[cfg] if isinstance(cfg, str) else cfg
    P = spec.P if P is None else P
    g = torch.Generator(device="cpu").manual_seed(seed)
    rn = lambda *s: torch.randn(*s, generator=g, dtype=torch.float32)
    ru = lambda *s: torch.rand(*s, generator=g, dtype=torch.float32)
    if spec.mean_kind == "normal":
        means = rn(P, 3) * spec.mean_args[0]
    else:
        ax, ay, sz = spec.mean_args
        means = torch.stack([(ru(P) * 2 - 1) * ax, (ru(P) * 2 - 1) * ay, rn(P) * sz], dim=1)
    scales = torch.exp(rn(P, 3) * spec.log_scale_sigma + spec.log_scale_mu)
    q = rn(P, 4)
    rotations = q / q.norm(dim=1, keepdim=True)
    opacities = torch.sigmoid(rn(P, 1) * 1.5)
    dc = rn(P, 1, 3)
    sh = rn(P, 15, 3) * 0.1
    w2c = torch.eye(4)
    w2c[2, 3] = -spec.eye_z  # eye at (0,0,eye_z) looking down +z
    cam = make_camera(spec.W, spec.H, spec.fx, spec.fx, w2c)
    out = dict(means3D=means.contiguous(), dc=dc, sh=sh, opacities=opacities, scales=scales,
               rotations=rotations.contiguous(), sh_degree=sh_degree, name=spec.name)
    out.update(cam)
    if device != "cpu":
        out = {k: (v.to(device) if isinstance(v, torch.Tensor) else v) for k, v in out.items()}
    return out


def raster_settings(mod, scene, bg=None, antialiasing=False, depth_threshold=0.0, debug=False,
                    scale_modifier=1.0, **extensions):
    """extensions: fields only this package's settings have (raw_activations, admm, f_count)."""
    dev = scene["means3D"].device
    if bg is None:
        bg = torch.zeros(3, dtype=torch.float32, device=dev)
    return mod.GaussianRasterizationSettings(
        image_height=int(scene["H"]), image_width=int(scene["W"]), tanfovx=scene["tanfovx"],
        tanfovy=scene["tanfovy"], bg=bg, scale_modifier=scale_modifier,
        viewmatrix=scene["viewmatrix"], projmatrix=scene["projmatrix"],
        sh_degree=int(scene["sh_degree"]), campos=scene["campos"], prefiltered=False, debug=debug,
        antialiasing=antialiasing, depth_threshold=depth_threshold, **extensions)


def render(mod, scene, params=None, **kw):
    """gaussian_render.render() with separate_sh=True: returns (image, radii, invdepth, screen_pts)."""
    p = scene if params is None else params
    rs = raster_settings(mod, scene, **kw)
    rasterizer = mod.GaussianRasterizer(raster_settings=rs)
    screen_space_points = torch.zeros_like(p["means3D"], requires_grad=True) + 0
    try:
        screen_space_points.retain_grad()
    except Exception:  # pylint: disable=W0702
        pass
    image, radii, depth = rasterizer(means3D=p["means3D"], means2D=screen_space_points, dc=p["dc"],
                                     shs=p["sh"], colors_precomp=None, opacities=p["opacities"],
                                     scales=p["scales"], rotations=p["rotations"], cov3D_precomp=None)
    return image, radii, depth, screen_space_points


# ---- ADMM block scenes (BASELINE.json configs 4 and 5) -------------------------------------
# A 4 x 2 grid of unit cells in xy; block b owns its cell expanded x1.4 about its centre
# (config/gaussian_splatting/urban3d_admm.yaml:18-20), so neighbouring blocks overlap by 0.4 and corner
# Gaussians have up to four owners.  The global cloud is uniform over the union of the expanded cells, which
# gives every block the same expected size; block b's camera hovers over its cell centre (nadir, height 3).
ADMM_SPECS = {"C4": dict(per_block=2_000_000, grid=(4, 2)), "C5": dict(per_block=4_000_000, grid=(4, 2))}
ADMM_EXPAND = 1.4
ADMM_DEPTH_THRESHOLD = 0.23             # urban3d_admm.yaml:86


def admm_layout(cfg, seed=0, per_block=None):
    """(xyz float32 [n_global, 3] on CPU, [global_indices int64 per block], boxes float64 [n_blocks, 2, 3])."""
    spec = ADMM_SPECS[cfg]
    gx, gy = spec["grid"]
    per_block = spec["per_block"] if per_block is None else per_block
    m = 0.5 * (ADMM_EXPAND - 1.0)
    n_global = int(round(per_block / (ADMM_EXPAND * ADMM_EXPAND) * (gx + 2 * m) * (gy + 2 * m)))
    g = torch.Generator(device="cpu").manual_seed(1000 + seed)
    xy = torch.rand(n_global, 2, generator=g, dtype=torch.float32)
    xyz = torch.stack([xy[:, 0] * (gx + 2 * m) - m, xy[:, 1] * (gy + 2 * m) - m,
                       torch.randn(n_global, generator=g, dtype=torch.float32) * 0.1], dim=1).contiguous()
    boxes, gidx = [], []
    h = 0.5 * ADMM_EXPAND
    for by in range(gy):
        for bx in range(gx):
            cx, cy = bx + 0.5, by + 0.5
            boxes.append([[cx - h, cy - h, -10.0], [cx + h, cy + h, 10.0]])
            inside = (xyz[:, 0] >= cx - h) & (xyz[:, 0] <= cx + h) & (xyz[:, 1] >= cy - h) & (xyz[:, 1] <= cy + h)
            gidx.append(torch.nonzero(inside).squeeze(1))
    return xyz, gidx, torch.tensor(boxes, dtype=torch.float64)


def make_admm_block(cfg, block, xyz_block, seed=0, device="cpu", W=1920, H=1080, fx=1500.0, sh_degree=3):
    """Scene dict of one block with RAW consensus parameters (pre-activation opacity / scaling / quaternion, the
    ADMM variables of slave_gaussian_trainer.py:81-97), distributions as C3."""
    gx, _ = ADMM_SPECS[cfg]["grid"]
    P = xyz_block.shape[0]
    g = torch.Generator(device="cpu").manual_seed(7919 * (seed + 1) + block)
    rn = lambda *s: torch.randn(*s, generator=g, dtype=torch.float32)
    out = dict(means3D=xyz_block.contiguous(), dc=rn(P, 1, 3), sh=rn(P, 15, 3) * 0.1,
               opacities=rn(P, 1) * 1.5, scales=rn(P, 3) * 0.7 + math.log(0.006), rotations=rn(P, 4),
               sh_degree=sh_degree, name=f"{cfg}/block{block}")
    w2c = torch.eye(4)
    w2c[0, 3], w2c[1, 3], w2c[2, 3] = -((block % gx) + 0.5), -((block // gx) + 0.5), 3.0
    out.update(make_camera(W, H, fx, fx, w2c))
    if device != "cpu":
        out = {k: (v.to(device) if isinstance(v, torch.Tensor) else v) for k, v in out.items()}
    return out
