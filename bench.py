#!/usr/bin/env python
"""bench.py — benchmarks of the DOGS rasterizer hot path on B200.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--cfg C3|C1|C2|C4|C5]

Metric (BASELINE.json): 3DGS forward+backward iterations/s at 1080p with 3 M Gaussians; ADMM block
training iterations/s at 1/2/4/8 B200.

A step = rasterizer forward + loss (0.8*L1 + 0.2*(1 - fused_ssim), gaussian_trainer.py:403-405) + backward
to every Gaussian parameter, on a seeded synthetic scene.

  N = 1   headline: config C3 (3 M Gaussians, 1920x1080).
          value  device-timed (CUDA events) iterations/s with scene, camera and target resident in HBM.
          e2e    the same step through the public drop-in API with the per-step HOST inputs of a training
                 iteration (camera matrices + target image, pinned) copied H2D and the loss read back D2H
                 inside the timed region (gaussian_trainer.py:338-359, :415-427).  The Gaussians are model
                 state and stay resident, as in the reference.
          configs   the same step on C1 and C2 (short runs) so every single-GPU config has a number in both arms.
          weak_admm the N > 1 step (below) on this one GPU, so the weak-scaling line has an N = 1 point
                    that contains the ADMM work.
  N > 1   one process per GPU, weak scaling: one C3-sized block per rank with RAW consensus parameters
          (activations fused into the rasterizer), the ADMM penalty of all 59 parameters added by the
          rasterizer backward every iteration, and one consensus round per timed region (i.e. every K
          iterations; the reference's interval is 200).  value = all ranks' iterations / max-over-ranks time.
  admm    (every run of this arm, any N) the SURVEY.md 8(e) protocol on config C4: 8 blocks x 2 M Gaussians
          in a 4 x 2 grid with x1.4 expansion, depth_threshold 0.23, blocks time-sliced 8/N per rank so the
          total work is the same at every N: efficiency = t_1 / (N * t_N).  `--cfg C4|C5` makes it the
          top-level metric (strong scaling).
  --impl reference   the UNMODIFIED reference CUDA extensions (oracle/_ref, built from /root/reference by
          oracle/build_ref.sh) through their own public API, same workload; its cpu_baseline field is the
          float32 PyTorch evaluation of the same math on the host cores (oracle/torch_baseline.py).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

METRIC = "3dgs_fwd_bwd_iters_per_s_1080p_3M"
UNIT = "iters/s"
PARAM_KEYS = ("means3D", "dc", "sh", "scales", "rotations", "opacities")     # the order of the ADMM groups


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region through NVML in-process
    (nvidia_ml_py).  Spawning `nvidia-smi` from the benchmark process stalls kernel launches for tens of
    milliseconds (fork of a CUDA process + driver locks), so it is only the fallback, used once
    before the region starts."""

    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20),
               ("sw_power_cap", 0x4))

    def __init__(self, index, period=0.05):
        self.index, self.period = index, period
        self.samples, self.reasons = [], set()
        self.sm_max = None
        self._stop = threading.Event()
        self.thread = None
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:  # pylint: disable=W0703
            self._smi_once()

    def _smi_once(self):
        try:
            out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits",
                                  "-i", str(self.index)], capture_output=True, text=True, timeout=10).stdout
            f = [float(x) for x in out.strip().split(",")]
            self.samples.append(f[0])
            self.sm_max = f[1]
        except Exception:  # pylint: disable=W0703
            pass

    def sample(self):
        """One NVML read (a few hundred microseconds of host time).  Called from the benchmark loop
        itself while the GPU still has queued work: a sampling thread or a background `nvidia-smi -lms`
        was measured to stall kernel launches for 20-140 ms per query."""
        n = self.nvml
        if n is None:
            return
        try:
            self.samples.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
            mask = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
            for name, bit in self.REASONS:
                if mask & bit:
                    self.reasons.add(name)
        except Exception:  # pylint: disable=W0703
            pass

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons), "samples": 0}
        sm = sorted(self.samples)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons),
                "samples": len(sm), "source": "nvml" if self.nvml else "nvidia-smi (before the region)"}


def load_modules(impl):
    if impl == "reference":
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import refutil
        if not refutil.ref_available():
            return None
        dgr, ssim, _ = refutil.ref_modules()
        return dgr, ssim.fused_ssim
    import dogs_b200.diff_gaussian_rasterization as dgr
    from dogs_b200.fused_ssim import fused_ssim
    return dgr, fused_ssim


# ---- CPU baselines (reported, not optimised; rank 0, N = 1) ----------------------------------------------------
def cpu_baseline_port(scene_cpu, n_tiles_sample=160, seed=0):
    """The scalar C oracle (kind 'port', 1 thread) on a bounded sample: preprocess + binning of ALL Gaussians,
    blend forward + backward on a random subset of tiles scaled to the full image."""
    import numpy as np
    from oracle import oracle as O
    s = {k: (v.numpy() if isinstance(v, torch.Tensor) else v) for k, v in scene_cpu.items()}
    W, H = int(s["W"]), int(s["H"])
    T = ((W + 15) // 16) * ((H + 15) // 16)
    rng = np.random.default_rng(seed)
    sel = np.zeros(T, np.uint8)
    sel[rng.choice(T, size=min(n_tiles_sample, T), replace=False)] = 1
    t0 = time.perf_counter()
    pre = O.preprocess(s)
    bins = O.binning(s, pre)
    t1 = time.perf_counter()
    img = O.render_fwd(s, pre, bins, tile_sel=sel)
    dL = rng.standard_normal((3, H, W)).astype(np.float32) / (3 * H * W)
    t2 = time.perf_counter()
    O.backward(s, pre, bins, img, dL, None, tile_sel=sel)   # includes preprocess backward over all P
    t3 = time.perf_counter()
    per_gauss = (t1 - t0)
    tile_fwd = (t2 - t1)
    tile_bwd = max((t3 - t2) - per_gauss, 0.0)
    scale = T / float(sel.sum())
    total = 2 * per_gauss + scale * (tile_fwd + tile_bwd)
    return {"value": 1.0 / total, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"all {s['means3D'].shape[0]} Gaussians preprocessed+binned+preprocess-bwd, blend fwd+bwd on "
                      f"{int(sel.sum())} of {T} tiles scaled x{scale:.1f}; {t3 - t0:.1f} s of CPU work",
            "est_s_per_iter": total}


def cpu_baseline(cfg):
    """north_star's CPU baseline: a float32 PyTorch evaluation of the same math (preprocess -> per-tile sort ->
    per-tile blend -> loss -> autograd backward) on all host cores, core count stated; the scalar C port is
    kept as a second field."""
    from dogs_b200 import scenes
    out = {}
    sc_cpu = scenes.make_scene(cfg, seed=0, device="cpu")
    try:
        from oracle import torch_baseline as TB
        out = TB.measure(sc_cpu, unit=UNIT)
    except Exception as ex:  # pylint: disable=W0703
        out = {"error": f"torch baseline: {ex}"}
    try:
        out["port"] = cpu_baseline_port(sc_cpu)
    except Exception as ex:  # pylint: disable=W0703
        out["port"] = {"error": str(ex)}
    if "value" not in out and "value" in out.get("port", {}):
        out.update({k: out["port"][k] for k in ("value", "unit", "cores", "kind", "sample")})
    return out


# ---- the training step -----------------------------------------------------------------------------------------
def to_raw(sc):
    """Activated scene -> raw consensus parameters (inverse sigmoid / log / un-normalised quaternion)."""
    raw = dict(sc)
    o = sc["opacities"].clamp(1e-4, 1 - 1e-4)
    raw["opacities"] = torch.log(o / (1 - o))
    raw["scales"] = torch.log(sc["scales"])
    return raw


class Trainee:
    """One model (scene block): leaf parameters in the model's own shapes, one training step on them."""

    def __init__(self, dgr, loss_of, sc, target, raw=False, admm_block=None, depth_threshold=0.0, optimizer=False):
        from dogs_b200 import scenes
        self.scenes, self.dgr, self.loss_of = scenes, dgr, loss_of
        self.sc, self.target = sc, target
        P = sc["means3D"].shape[0]
        self.P = P
        # leaf parameters in the model's own (ADMM) shapes: f_dc [P,3], f_rest [P,45]; views feed the rasterizer
        self.leaves = {"means3D": sc["means3D"].clone(), "dc": sc["dc"].reshape(P, 3).clone(),
                       "sh": sc["sh"].reshape(P, 45).clone(), "opacities": sc["opacities"].clone(),
                       "scales": sc["scales"].clone(), "rotations": sc["rotations"].clone()}
        for v in self.leaves.values():
            v.requires_grad_(True)
        self.kw = {}
        if raw:
            self.kw["raw_activations"] = True
        if depth_threshold:
            self.kw["depth_threshold"] = depth_threshold
        self.blk = admm_block
        # SparseGaussianAdam over the six tensors, step(radii > 0, P) as at gaussian_trainer.py:474-476.  The
        # learning rates are 0: the update runs in full (moments and parameters are read and written) but the scene
        # stays the seeded one — against a random target image real rates inflate the Gaussians within a few steps
        # and every later step renders a different, much heavier scene.
        self.opt = None
        if optimizer:
            self.opt = dgr.SparseGaussianAdam([{"params": [self.leaves[k]], "lr": 0.0, "name": k} for k in PARAM_KEYS],
                                              lr=0.0, eps=1e-15)

    def x(self):
        return [self.leaves[k] for k in PARAM_KEYS]

    def attach(self, blk):
        self.blk = blk
        self.kw["admm"] = blk

    def step(self, scene=None, target=None):
        lv = self.leaves
        for v in lv.values():
            v.grad = None
        views = dict(lv, dc=lv["dc"].view(self.P, 1, 3), sh=lv["sh"].view(self.P, 15, 3))
        img, radii, _, _ = self.scenes.render(self.dgr, self.sc if scene is None else scene, views, **self.kw)
        loss = self.loss_of(img, self.target if target is None else target)
        loss.backward()
        if self.blk is not None:
            self.blk.add_penalty_grads_()
        if self.opt is not None:
            self.opt.step(radii > 0, self.P)
        return loss


def virtual_topology(P, dev, frac=0.245):
    """One GPU, no peers: the first and last `frac` of the block are marked as shared with a (virtual) second
    owner, so the per-iteration penalty and the consensus kernels do the work of the N > 1 path."""
    from dogs_b200 import admm as A
    ov = int(frac * P)
    shared = torch.zeros(P, dtype=torch.bool)
    shared[:ov] = True
    shared[P - ov:] = True
    plan = A.BlockPlan(0, P, P)
    plan.S = int(shared.sum())
    plan.shared_row = torch.where(shared, torch.cumsum(shared.long(), 0) - 1, torch.full((P,), -1)).to(torch.int32)
    plan.max_addends = 1
    plan.addends = torch.full((plan.S, 1), A.SELF_ADDEND - (1 << 32), dtype=torch.int32)
    plan.count = torch.full((plan.S,), 2, dtype=torch.uint8)
    plan.pack_src = torch.nonzero(shared).squeeze(1).to(torch.int32)     # packed (as if sent), never received
    plan.send_rows = plan.S
    topo = A.Topology()
    topo.n_global = P
    topo.send_rows_total = plan.S
    topo.plans[0] = plan
    return topo.to(dev)


def timed_steps(step, K, barrier, after=None):
    """K steps (+ `after`, e.g. one consensus round) between two events on the current stream.
    Returns (total ms, sorted per-step ms)."""
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    marks = [ev0]
    for _ in range(K):
        step()
        m = torch.cuda.Event(enable_timing=True)
        m.record()
        marks.append(m)
    if after is not None:
        after()
    ev1.record()
    return ev0, ev1, marks


def admm_cpu_baseline(gidx, n_global, blocks=(0, 1)):
    """CPU baseline of the consensus round: the reference's master / slave arithmetic — scatter-add
    of every owner's 59 parameters, division by the owner count, dual update (master_gaussian_trainer.py:538-555,
    slave_gaussian_trainer.py:99-121) — as restated in oracle/oracle.py (numpy, one core), on a bounded sample of
    the C4 blocks, scaled to all eight."""
    import numpy as np
    from oracle import oracle as O
    rng = np.random.default_rng(0)
    idx = [gidx[b].numpy() for b in blocks]
    xs = [[rng.standard_normal((len(i), sz), dtype=np.float32) for sz in O.GROUP_SIZES] for i in idx]
    us = [[np.zeros_like(a) for a in x] for x in xs]
    t0 = time.perf_counter()
    with np.errstate(invalid="ignore", divide="ignore"):     # Gaussians of the blocks outside the sample have no owner
        z, _ = O.admm_consensus(xs, idx, n_global)
    for x, u, i in zip(xs, us, idx):
        O.admm_dual_update(x, u, [zg[i] for zg in z])
    dt = time.perf_counter() - t0
    est = dt * 8.0 / len(blocks)
    return {"value": 1.0 / est, "unit": "consensus rounds/s", "cores": 1, "kind": "port",
            "sample": f"{len(blocks)} of 8 blocks ({sum(len(i) for i in idx)} Gaussians): scatter-add, average, dual update "
                      f"in numpy ({dt:.1f} s), scaled x{8 // len(blocks)}", "est_s_per_round": est}


def admm_protocol(cfg, dgr, loss_of, K, warm, rank, world, dev, distributed, cpu_baseline_leg=False):
    """8 blocks (C4: 2 M, C5: 4 M Gaussians each), 4 x 2 grid, x1.4 expansion, time-sliced over
    the ranks.  One ADMM iteration = every block runs one training step (render of its own camera with the
    fused penalty, loss, backward); one consensus round closes the timed region."""
    from dogs_b200 import admm as A, scenes
    nb = 8
    if nb % world:
        return {"skipped": f"8 blocks do not divide over {world} ranks"}
    per_rank = nb // world
    mine = list(range(rank * per_rank, (rank + 1) * per_rank))
    xyz, gidx, _ = scenes.admm_layout(cfg)
    n_global = xyz.shape[0]
    torch.manual_seed(4321)
    H, W = 1080, 1920
    target = torch.rand(3, H, W, device=dev)
    models = {}
    for b in mine:
        sc = scenes.make_admm_block(cfg, b, xyz[gidx[b]], device=dev, W=W, H=H)
        models[b] = Trainee(dgr, loss_of, sc, target, raw=True, depth_threshold=scenes.ADMM_DEPTH_THRESHOLD)
    topo = A.build_topology({b: gidx[b].to(dev) for b in mine}, n_global)
    grp = A.ConsensusGroup({b: models[b].x() for b in mine}, topo)
    for b in mine:
        models[b].attach(grp.blocks[b].fuse_into_rasterizer(True, all_groups=True))
    del xyz
    check = None
    if distributed:
        # untimed self-check of the transport: the same round through batched NCCL send/recv must give the same z
        ref = A.ConsensusGroup({b: [t.detach().clone() for t in models[b].x()] for b in mine}, topo, p2p=False)
        chk = A.ConsensusGroup({b: [t.detach().clone() for t in models[b].x()] for b in mine}, topo)
        ref.consensus(adapt=False)
        chk.consensus(adapt=False)
        dz = torch.stack([(p_ - q_).abs().max() for b in mine for p_, q_ in zip(ref.blocks[b].z, chk.blocks[b].z)]).max()
        dist.all_reduce(dz, op=dist.ReduceOp.MAX)
        check = {"transport": ("peer memory: the pack kernel writes the co-owners' receive buffers (torch symmetric memory)"
                               if chk.p2p is not None else "NCCL send/recv"),
                 "peer_memory_error": chk.p2p_error, "max_abs_dz_vs_nccl_send_recv": float(dz.item())}
        del ref, chk
        torch.cuda.empty_cache()

    def iteration():
        for b in mine:
            models[b].step()

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warm):
        iteration()
    grp.consensus()
    ev0, ev1, marks = timed_steps(iteration, K, barrier, after=grp.consensus)
    barrier()
    ms = ev0.elapsed_time(ev1)
    it_seq = [marks[i].elapsed_time(marks[i + 1]) for i in range(K)]
    it_ms = sorted(it_seq)
    rounds = []
    for _ in range(3):
        barrier()
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        grp.consensus(adapt=False)
        b_.record()
        barrier()
        rounds.append(a.elapsed_time(b_))
    t = torch.tensor([ms, sorted(rounds)[1]], dtype=torch.float64, device=dev)
    if distributed:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, round_ms = float(t[0]), float(t[1])
    sizes = [int(g.numel()) for g in gidx]
    shared = sum(int(topo.plans[b].S) for b in mine)
    out = {"cfg": cfg, "protocol": "SURVEY 8(e): 8 blocks time-sliced over the ranks, constant total work (strong scaling)",
           "blocks": nb, "blocks_per_rank": per_rank, "n_gpus": world, "gaussians_per_block": sizes, "n_global": n_global,
           "image": [W, H], "depth_threshold": scenes.ADMM_DEPTH_THRESHOLD, "raw_activations": True,
           "penalty": "all 59 parameters, fused into the rasterizer backward, every iteration",
           "iters": K, "consensus_rounds_in_timed_region": 1, "ms_total": ms, "ms_per_iter": ms / K,
           "iters_per_s": K / (ms / 1e3), "block_iters_per_s": nb * K / (ms / 1e3),
           "iter_ms_rank0": {"min": it_ms[0], "median": it_ms[K // 2], "max": it_ms[-1], "in_order": [round(t, 3) for t in it_seq]},
           "consensus_round_ms": round_ms, "consensus_interval_reference": 200,
           "iters_per_s_at_interval_200": 1e3 / (it_ms[K // 2] + round_ms / 200.0),
           "shared_rows_rank0": shared, "exchange_bytes_rank0": {"send_buffer": int(topo.send_rows_total) * 236,
                                                                 "recv_buffer": int(topo.recv_rows_total) * 236},
           "hbm_gb_rank0": torch.cuda.max_memory_allocated(dev) / 1e9,
           "exchange": check if check is not None else {"transport": "none: all blocks on one GPU read each other's send rows in place"}}
    if cpu_baseline_leg and rank == 0:
        try:
            out["cpu_baseline"] = admm_cpu_baseline(gidx, n_global)
            out["cpu_baseline"]["gpu_rounds_per_s"] = 1e3 / round_ms
        except Exception as ex:  # pylint: disable=W0703
            out["cpu_baseline"] = {"error": repr(ex)}
    del grp, models
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cfg", default="C3")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-admm", action="store_true", help="skip the C4 ADMM protocol sub-object")
    ap.add_argument("--no-extra", action="store_true", help="skip the C1 / C2 / weak_admm extras at N = 1")
    ap.add_argument("--force-admm", action="store_true",
                    help="profiling aid: make the N>1 step (raw parameters + fused ADMM penalty) the top-level workload on one GPU")
    ap.add_argument("--unfused-loss", action="store_true",
                    help="ours: evaluate the loss with torch ops around fused_ssim, as the reference's trainer does")
    ap.add_argument("--no-clocks", action="store_true", help="do not query NVML inside the timed region")
    ap.add_argument("--no-roofline", action="store_true", help="skip the separately timed blend_bwd launches (ncu launch lists)")
    args = ap.parse_args()
    K, Wm = args.steps, max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference" and rank != 0:
        return 0                                    # rank 0 alone runs the reference arm
    if not torch.cuda.is_available():
        print(json.dumps({"impl": args.impl, "unavailable": "no CUDA device (the hot path has no CPU fallback)"}))
        return 0
    torch.cuda.set_device(local_rank)
    dev = f"cuda:{local_rank}"
    ours = args.impl == "ours"
    distributed = world > 1 and ours
    if distributed:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(dev))

    mods = load_modules(args.impl)
    if mods is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref not built (run oracle/build_ref.sh where /root/reference exists)"}))
        return 0
    dgr, fused_ssim = mods
    from dogs_b200 import scenes

    fused_loss = None
    if ours and not args.unfused_loss:
        from dogs_b200.fused_ssim import fused_l1_ssim_loss as fused_loss   # same loss, 2 launches

    def loss_of(img, tgt):
        if fused_loss is not None:
            return fused_loss(img, tgt, 0.2)
        l1 = (img - tgt).abs().mean()
        return 0.8 * l1 + 0.2 * (1.0 - fused_ssim(img[None], tgt[None]))

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    lib = None
    if ours:
        from dogs_b200 import _lib
        lib = _lib.lib()

    def allmax(ms):
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if distributed:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- top-level = the ADMM protocol itself (--cfg C4 | C5) ---------------------------------------------------
    if args.cfg in ("C4", "C5"):
        if not ours:
            print(json.dumps({"impl": "reference", "unavailable": "the ADMM protocol bench exists for this arm only"}))
            return 0
        n0 = lib.dogs_launch_count()
        clk = ClockSampler(local_rank)
        a = admm_protocol(args.cfg, dgr, loss_of, K, Wm, rank, world, dev, distributed)
        clk.sample()
        if rank == 0:
            print(json.dumps({"metric": f"admm_iters_per_s_{args.cfg}_8blocks", "value": a.get("iters_per_s"), "unit": UNIT,
                              "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": a.get("ms_per_iter"),
                              "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                              "data": "synthetic", "config": {"workload": f"{args.cfg}: 8 ADMM blocks, see admm"},
                              "gpu_launches": int(lib.dogs_launch_count() - n0), "clocks": clk.summary(), "admm": a}), flush=True)
        if distributed:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    # ---- the block of this rank -------------------------------------------------------------------------------
    with_admm = distributed or args.force_admm
    sc = scenes.make_scene(args.cfg, seed=rank, device=dev)
    torch.manual_seed(1234 + rank)
    target = torch.rand(3, sc["H"], sc["W"], device=dev)
    P, H, Wd = sc["means3D"].shape[0], sc["H"], sc["W"]

    def make_admm_trainee():
        """The N > 1 step: raw consensus parameters, all 59 penalty groups fused into the rasterizer backward."""
        from dogs_b200 import admm as A
        tr = Trainee(dgr, loss_of, to_raw(sc), target, raw=True)
        if distributed:
            # 1-D chain of blocks, 24.5 % of a block shared with each neighbour (~49 % shared)
            ov = int(0.245 * P)
            stride = P - ov
            gidx = torch.arange(P, dtype=torch.int64, device=dev) + rank * stride
            topo = A.build_plan(gidx, stride * (world - 1) + P)
        else:
            topo = virtual_topology(P, dev)
        blk = A.AdmmBlock(tr.x(), topo)
        tr.attach(blk.fuse_into_rasterizer(True, all_groups=True))
        return tr, blk

    blk = None
    if with_admm:
        model, blk = make_admm_trainee()
    else:
        model = Trainee(dgr, loss_of, sc, target)
    step = model.step
    consensus = (lambda: blk.consensus()) if blk is not None else None

    for _ in range(Wm):
        step()
    if consensus:
        consensus()
    barrier()

    # ---- value: inputs resident in HBM, device-timed --------------------------------------------------------------
    n0 = lib.dogs_launch_count() if lib else 0
    clk = ClockSampler(local_rank)
    ev0, ev1, marks = timed_steps(step, K, barrier, after=consensus)
    # Clocks / throttle reasons under load: the host has enqueued the whole region, the GPU is still executing
    # its tail (ev1 not reached yet).  Querying NVML between launches instead was measured to stall the
    # launching thread for 10-25 ms per query.
    if not args.no_clocks:
        for _ in range(3):
            clk.sample()
    barrier()
    ms_total = allmax(ev0.elapsed_time(ev1))
    per_step = sorted(marks[i].elapsed_time(marks[i + 1]) for i in range(K))
    launches = (lib.dogs_launch_count() - n0) if lib else 0
    nranks = world if ours else 1
    value = nranks * K / (ms_total / 1e3)

    # ---- e2e: host inputs of a training iteration H2D, loss D2H, every step -------------------------------------
    e2e = None
    if not args.no_e2e:
        host_target = torch.empty(3, H, Wd, dtype=torch.float32).pin_memory()
        host_target.copy_(target.cpu())
        host_cam = {k: sc[k].cpu().pin_memory() for k in ("viewmatrix", "projmatrix", "campos")}
        h2d = host_target.numel() * 4 + sum(v.numel() * 4 for v in host_cam.values())

        # Per-step host inputs are double-buffered on a copy stream (the H2D copy of step i+1 overlaps the
        # compute of step i) and the loss is read back with a one-step lag through pinned memory, the way
        # an asynchronous data loader / logger drives a trainer.  Every step still copies its own target
        # image + camera H2D and its own loss D2H inside the timed region.
        copy_stream = torch.cuda.Stream(device=dev)
        dev_target = [torch.empty_like(target) for _ in range(2)]
        dev_cam = [{k: torch.empty_like(sc[k]) for k in host_cam} for _ in range(2)]
        ready = [torch.cuda.Event() for _ in range(2)]
        consumed = [torch.cuda.Event() for _ in range(2)]
        loss_host = [torch.zeros(1, dtype=torch.float32).pin_memory() for _ in range(2)]
        loss_ev = [torch.cuda.Event() for _ in range(2)]
        computed = [torch.cuda.Event() for _ in range(2)]
        losses = []

        def prefetch(i):
            b = i % 2
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(consumed[b])           # the step that last used buffer b is finished
                dev_target[b].copy_(host_target, non_blocking=True)
                for k, v in host_cam.items():
                    dev_cam[b][k].copy_(v, non_blocking=True)
                ready[b].record(copy_stream)

        def e2e_loop(n, marks=None):
            cur = torch.cuda.current_stream(dev)
            for b in range(2):
                consumed[b].record(cur)
            prefetch(0)
            for i in range(n):
                b = i % 2
                if i + 1 < n:
                    prefetch(i + 1)
                cur.wait_event(ready[b])
                scene = dict(model.sc)
                scene.update(dev_cam[b])
                loss = step(scene, dev_target[b])
                consumed[b].record(cur)
                # D2H of this step's loss (4 B) on the copy stream, so that no copy node sits between two steps
                # of the compute stream
                computed[b].record(cur)
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(computed[b])
                    lv = loss.detach().reshape(1)
                    lv.record_stream(copy_stream)
                    loss_host[b].copy_(lv, non_blocking=True)
                    loss_ev[b].record(copy_stream)
                if marks is not None:
                    m_ = torch.cuda.Event(enable_timing=True)
                    m_.record(cur)
                    marks.append(m_)
                if i >= 1:                                     # read the previous step's loss on the host
                    loss_ev[1 - b].synchronize()
                    losses.append(float(loss_host[1 - b][0]))
            loss_ev[(n - 1) % 2].synchronize()
            losses.append(float(loss_host[(n - 1) % 2][0]))

        e2e_loop(3)
        barrier()
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        e2e_marks = [a]
        e2e_loop(K, e2e_marks)
        if consensus:
            consensus()
        b_.record()
        barrier()
        e2e_steps = [e2e_marks[i].elapsed_time(e2e_marks[i + 1]) for i in range(len(e2e_marks) - 1)]
        e2e = {"value": nranks * K / (allmax(a.elapsed_time(b_)) / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": 4,
               "step_ms": {"first": round(e2e_steps[0], 3), "median": round(sorted(e2e_steps)[len(e2e_steps) // 2], 3),
                           "max": round(max(e2e_steps), 3), "tail_ms": round(e2e_marks[-1].elapsed_time(b_), 3)}}

    # ---- roofline of the dominant kernel (blend backward), timed live with CUDA events ---------------------------
    roofline, stats = None, {}
    if ours and rank == 0 and not args.no_roofline:
        roofline, stats = blend_bwd_roofline(sc, dev)

    # ---- extras at N = 1: the other single-GPU configs, and the N > 1 step on one GPU --------------------------
    extra = {}
    if world == 1 and not args.no_extra and not with_admm:
        del model
        torch.cuda.empty_cache()
        cfgs = {}
        for c in ("C1", "C2"):
            s2 = scenes.make_scene(c, seed=0, device=dev)
            t2 = torch.rand(3, s2["H"], s2["W"], device=dev)
            m2 = Trainee(dgr, loss_of, s2, t2)
            for _ in range(Wm):
                m2.step()
            e0, e1, mk = timed_steps(m2.step, K, barrier)
            barrier()
            cfgs[c] = {"value": K / (e0.elapsed_time(e1) / 1e3), "unit": UNIT, "ms_per_step": e0.elapsed_time(e1) / K,
                       "gaussians": s2["means3D"].shape[0], "image": [s2["W"], s2["H"]]}
            del m2, s2, t2
        extra["configs"] = cfgs
        # the same steps replayed as one CUDA graph each (dogs_b200.graphs.GraphedStep: forward in the library's
        # deferred-check mode, verified after every replay): what is left when the host no longer paces the step
        if ours:
            try:
                from dogs_b200.graphs import GraphedStep
                gcfg = {}
                for c in ("C1", "C2"):
                    s2 = scenes.make_scene(c, seed=0, device=dev)
                    t2 = torch.rand(3, s2["H"], s2["W"], device=dev)
                    m2 = Trainee(dgr, loss_of, s2, t2)
                    gs = GraphedStep(m2.step, warmup=Wm)
                    for _ in range(2):
                        gs()
                    e0, e1, mk = timed_steps(gs, K, barrier)
                    barrier()
                    gcfg[c] = {"value": K / (e0.elapsed_time(e1) / 1e3), "unit": UNIT, "ms_per_step": e0.elapsed_time(e1) / K,
                               "recaptures": gs.recaptures}
                    del gs, m2, s2, t2
                extra["graphed"] = gcfg
            except Exception as ex:  # pylint: disable=W0703
                extra["graphed"] = {"error": repr(ex)}
            torch.cuda.empty_cache()
        # the same C3 step followed by the sparse Adam update of all six tensors (SURVEY.md 8f.1,
        # gaussian_trainer.py:474-476): six adamUpdate launches in the reference, one here
        mo = Trainee(dgr, loss_of, sc, target, optimizer=True)
        for _ in range(Wm):
            mo.step()
        e0, e1, mk = timed_steps(mo.step, K, barrier)
        barrier()
        extra["with_optimizer"] = {"value": K / (e0.elapsed_time(e1) / 1e3), "unit": UNIT, "ms_per_step": e0.elapsed_time(e1) / K,
                                   "workload": f"{args.cfg} step + SparseGaussianAdam.step(radii > 0) over the six parameter tensors"}
        del mo
        torch.cuda.empty_cache()
        if ours:
            m3, b3 = make_admm_trainee()
            for _ in range(Wm):
                m3.step()
            b3.consensus()
            e0, e1, mk = timed_steps(m3.step, K, barrier, after=lambda: b3.consensus())
            barrier()
            ps = sorted(mk[i].elapsed_time(mk[i + 1]) for i in range(K))
            extra["weak_admm"] = {"value": K / (e0.elapsed_time(e1) / 1e3), "unit": UNIT, "n_gpus": 1,
                                  "ms_per_step": e0.elapsed_time(e1) / K, "step_ms_median": ps[K // 2],
                                  "workload": "the N > 1 step on one GPU: raw parameters, 59-parameter ADMM penalty fused "
                                              f"into the backward, 1 consensus round per {K} iterations (virtual neighbours)"}
            del m3, b3
        torch.cuda.empty_cache()

    admm_obj = None
    if ours and not args.no_admm and not args.force_admm:
        try:
            admm_obj = admm_protocol("C4", dgr, loss_of, max(K // 2, 5), 6, rank, world, dev, distributed,
                                     cpu_baseline_leg=(world == 1 and not args.no_cpu_baseline))
        except Exception as ex:  # pylint: disable=W0703
            admm_obj = {"error": repr(ex)}

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": nranks,
            "steps": K, "warmup": Wm, "ms_per_step": ms_total / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.cfg}: {P} Gaussians/block, {Wd}x{H}, SH degree 3, rasterize fwd + "
                                   f"0.8*L1+0.2*(1-SSIM) + bwd" + (", raw parameters (activations fused), ADMM penalty of all 59 "
                                   f"parameters per iteration + 1 consensus round per {K} iterations, one block per GPU"
                                   if with_admm else ""),
                       "l2_policy": "inputs (708 MB of Gaussian parameters + 10 M instance lists) exceed the 126 MB L2",
                       "loss": ("fused_l1_ssim_loss (one launch forward, one backward)" if fused_loss is not None
                                else "torch L1 ops + fused_ssim"),
                       "blocks": nranks},
            "gpu_launches": int(launches), "clocks": clk.summary(),
            "step_ms": {"min": per_step[0], "median": per_step[K // 2], "max": per_step[-1]},
        }
        if not ours:
            out["impl"] = "reference"
            out["reference_kind"] = "unmodified reference CUDA extensions (oracle/_ref) on the same GPU"
        if e2e is not None:
            out["e2e"] = e2e
        if roofline is not None:
            out["roofline"] = roofline
            out["scene_stats"] = stats
        out.update(extra)
        if admm_obj is not None:
            out["admm"] = admm_obj
        if not args.no_cpu_baseline and (world == 1 or not ours):
            out["cpu_baseline"] = cpu_baseline(args.cfg)
        print(json.dumps(out), flush=True)
    if distributed:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def blend_bwd_roofline(sc, dev, reps=10):
    """Times blend_bwd_kernel and preprocess_bwd_kernel alone with CUDA events on the launching stream
    (dogs_stage_times of the last dgr_backward) on a prepared forward state.  The kernel is instruction-issue
    bound (no stage is a dense contraction), so `bound` says "issue" and `frac` is the issue-slot
    fraction; the HBM figure the contract asks for is reported next to it with the algorithmic bytes the
    kernel must move:
      44 B per tile instance (4 B id + 40 B record) + 44 B per pixel (n_contrib, final_T, colour copy,
      inverse depth, dL/dC, dL/dD) + 40 B per instance of gradient flush."""
    import ctypes as C
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import refutil as U
    from dogs_b200 import _lib as L
    o = U.ours_forward_raw(sc)
    R = o["num_rendered"]
    H, Wd = sc["H"], sc["W"]
    N = H * Wd
    sum_nc = int(o["n_contrib"].long().sum().item())
    stats = {"num_rendered": int(R), "visible": int((o["radii"] > 0).sum().item()),
             "sum_n_contrib": sum_nc, "pixels": N}
    al, keep = o["_keep"]
    from dogs_b200.diff_gaussian_rasterization import _params, GaussianRasterizationSettings
    P = sc["means3D"].shape[0]
    bg = torch.zeros(3, device=dev)
    rs = GaussianRasterizationSettings(H, Wd, sc["tanfovx"], sc["tanfovy"], bg, 1.0, sc["viewmatrix"],
                                       sc["projmatrix"], 3, sc["campos"], False, False, False, 0.0)
    prm, keep2 = _params(rs, P, 15, dev)
    dLc = torch.randn(3, H, Wd, device=dev) / N
    new = lambda *s: torch.empty(s, device=dev)
    outs = [new(P, 3), new(P, 3), new(P, 3), new(P, 45), None, new(P, 1), new(P, 3), new(P, 4), None]
    scratch = L.Allocator(dev)
    stage = (C.c_float * 2)()
    best = [1e30, 1e30]
    L.lib().dogs_stage_timing(1)
    try:
        for _ in range(reps):
            L.check(L.lib().dgr_backward(C.byref(prm), R, sc["means3D"].data_ptr(), sc["dc"].data_ptr(), sc["sh"].data_ptr(), None,
                                         sc["opacities"].data_ptr(), sc["scales"].data_ptr(), sc["rotations"].data_ptr(), None,
                                         o["radii"].data_ptr(), al[0].last.data_ptr(), al[1].last.data_ptr(), al[2].last.data_ptr(),
                                         dLc.data_ptr(), None, *[L.ptr(t) for t in outs], scratch.cb, None,
                                         L.stream_ptr(dev)))
            L.check(L.lib().dogs_stage_times(stage, 2))
            best = [min(best[0], stage[0]), min(best[1], stage[1])]
            scratch.buffers.clear()
    finally:
        L.lib().dogs_stage_timing(0)
    peak, how = peaks()
    alg_bytes = 44.0 * R + 44.0 * N + 40.0 * R
    achieved = alg_bytes / (best[0] * 1e-3) / 1e9
    traffic, issue = None, None
    tpath = os.path.join(ROOT, "profiles", "r2_traffic.json")
    if not os.path.exists(tpath):
        tpath = os.path.join(ROOT, "profiles", "r1_traffic.json")
    if os.path.exists(tpath) and sc["name"] == "C3":
        with open(tpath) as f:
            t = json.load(f).get("blend_bwd_kernel", {})
        traffic = t.get("dram_bytes")
        if t.get("warp_instructions"):
            # issue-slot roofline: 148 SMs x 4 schedulers x 1 warp instruction / clock at the max SM clock
            issue_peak = 148 * 4 * 1.965e9
            issue = {"warp_instructions": t["warp_instructions"],
                     "achieved_ginst_s": t["warp_instructions"] / (best[0] * 1e-3) / 1e9,
                     "peak_ginst_s": issue_peak / 1e9,
                     "frac": t["warp_instructions"] / (best[0] * 1e-3) / issue_peak,
                     "source": f"ncu smsp__inst_executed.sum of the committed capture ({os.path.basename(tpath)}) / live kernel time"}
    roof = {"kernel": "blend_bwd_kernel", "bound": "issue" if issue else "hbm",
            "achieved": issue["achieved_ginst_s"] if issue else achieved,
            "peak": issue["peak_ginst_s"] if issue else peak,
            "unit": "Gwarp-inst/s" if issue else "GB/s",
            "frac": issue["frac"] if issue else achieved / peak,
            "hbm": {"achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "algorithmic_bytes": alg_bytes, "peak_source": how},
            "traffic": traffic, "kernel_ms": best[0], "issue": issue,
            "note": "instruction-issue bound, not HBM bound (no stage is a dense contraction): "
                    f"{sum_nc} visited (pixel,splat) pairs; preprocess_bwd_kernel {best[1]:.3f} ms"}
    return roof, stats


if __name__ == "__main__":
    sys.exit(main())
