"""Host-side cost of one training step: wall time per step with the GPU kept busy vs the device time, and the
top host functions (cProfile):  python tools/gpu_cpu_overhead.py C1 [steps]"""
import cProfile, io, os, pstats, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B
from dogs_b200 import scenes
import dogs_b200.diff_gaussian_rasterization as dgr
from dogs_b200.fused_ssim import fused_l1_ssim_loss as fused_loss
cfg = sys.argv[1] if len(sys.argv) > 1 else "C1"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
dev = "cuda:0"
sc = scenes.make_scene(cfg, 0, dev)
target = torch.rand(3, sc["H"], sc["W"], device=dev)
loss_of = lambda img, tgt: fused_loss(img, tgt, 0.2)
tr = B.Trainee(dgr, loss_of, sc, target)
for _ in range(10):
    tr.step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
for _ in range(steps):
    tr.step()
e1.record(); t_issue = time.perf_counter() - t0
torch.cuda.synchronize(); t_wall = time.perf_counter() - t0
print(f"{cfg}: host issue {1e3 * t_issue / steps:.3f} ms/step, wall {1e3 * t_wall / steps:.3f} ms/step, device span {e0.elapsed_time(e1) / steps:.3f} ms/step")
pr = cProfile.Profile(); pr.enable()
for _ in range(steps):
    tr.step()
torch.cuda.synchronize(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(18); print(s.getvalue()[:3500])
