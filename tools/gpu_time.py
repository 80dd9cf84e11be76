"""Times reference vs ours rasterizer fwd / fwd+bwd with CUDA events: python tools/gpu_time.py C3 [iters]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import refutil as U  # noqa: E402
from dogs_b200 import scenes  # noqa: E402
import dogs_b200.diff_gaussian_rasterization as ours_mod  # noqa: E402


def timeit(fn, iters, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], ts[0], ts[-1]


def main():
    cfg = sys.argv[1] if len(sys.argv) > 1 else "C1"
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    which = sys.argv[3] if len(sys.argv) > 3 else "both"
    dev = "cuda:0"
    sc = scenes.make_scene(cfg, 0, dev)
    H, W = sc["H"], sc["W"]
    target = torch.rand(3, H, W, device=dev)
    rep = {"cfg": cfg}
    mods = []
    if which in ("both", "ref") and U.ref_available():
        mods.append(("ref", U.ref_modules()[0]))
    if which in ("both", "ours"):
        mods.append(("ours", ours_mod))
    for name, mod in mods:
        p = {k: sc[k].clone().requires_grad_(True) for k in ("means3D", "dc", "sh", "opacities", "scales", "rotations")}

        def fwd():
            with torch.no_grad():
                scenes.render(mod, sc, p)

        def fwdbwd():
            for v in p.values():
                v.grad = None
            img, _, _, _ = scenes.render(mod, sc, p)
            loss = (img - target).abs().mean()
            loss.backward()

        rep[name + "_fwd_ms"] = timeit(fwd, iters)
        rep[name + "_fwdbwd_ms"] = timeit(fwdbwd, iters)
        print(name, rep[name + "_fwd_ms"], rep[name + "_fwdbwd_ms"], flush=True)
    print(json.dumps(rep))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"time_{cfg}_{which}.json"), "w") as f:
        json.dump(rep, f)


if __name__ == "__main__":
    main()
