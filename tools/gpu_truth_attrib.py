"""Where does the float32 backward lose accuracy at full size?  Splits the error against the FP64 truth into
(a) the per-Gaussian accumulators produced by the blend backward and (b) the per-Gaussian chain rule:
  ours               float32 blend  -> float32 preprocess        (the product path)
  acc32 -> pre64     float32 accumulators of the product blend fed to the FP64 preprocess
    python tools/gpu_truth_attrib.py C3"""
import ctypes as C, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import refutil as U
from dogs_b200 import scenes, _lib as L
from dogs_b200.diff_gaussian_rasterization import _params, GaussianRasterizationSettings
cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
DEV = "cuda:0"
keys = ("means3D", "means2D", "dc", "sh", "opacities", "scales", "rotations")
sc = scenes.make_scene(cfg, 0, DEV)
P, H, W = sc["means3D"].shape[0], sc["H"], sc["W"]
g = torch.Generator(device="cpu").manual_seed(1)
dLc = torch.randn(3, H, W, generator=g).to(DEV); dLd = torch.randn(1, H, W, generator=g).to(DEV)
o = U.ours_forward_raw(sc)
truth = U.truth_backward(sc, o, dLc, dLd)
# product backward through the C-ABI with our own scratch allocator: the accumulators stay readable
al, keep = o["_keep"]
bg = torch.zeros(3, device=DEV)
rs = GaussianRasterizationSettings(H, W, sc["tanfovx"], sc["tanfovy"], bg, 1.0, sc["viewmatrix"], sc["projmatrix"], 3,
                                   sc["campos"], False, False, False, 0.0)
prm, keep2 = _params(rs, P, 15, DEV)
new = lambda *s: torch.empty(s, device=DEV)
outs = [new(P, 3), new(P, 3), new(P, 3), new(P, 45), None, new(P, 1), new(P, 3), new(P, 4), None]
scratch = L.Allocator(DEV)
L.check(L.lib().dgr_backward(C.byref(prm), o["num_rendered"], sc["means3D"].data_ptr(), sc["dc"].data_ptr(), sc["sh"].data_ptr(), None,
                             sc["opacities"].data_ptr(), sc["scales"].data_ptr(), sc["rotations"].data_ptr(), None,
                             o["radii"].data_ptr(), al[0].last.data_ptr(), al[1].last.data_ptr(), al[2].last.data_ptr(),
                             dLc.data_ptr(), dLd.data_ptr(), *[L.ptr(t) for t in outs], scratch.cb, None, L.stream_ptr(DEV)))
torch.cuda.synchronize()
acc32 = scratch.buffers[0][:P * 48].view(torch.float32).view(P, 12)[:, :10].double().contiguous()
ours = dict(means3D=outs[0], means2D=outs[1], dc=outs[2], sh=outs[3], opacities=outs[5], scales=outs[6], rotations=outs[7])
# truth accumulators: recompute (truth_backward does not return them)
T = C.CDLL(os.path.join(ROOT, "oracle", "libtruth.so"))
acc64 = torch.zeros(P, 10, dtype=torch.float64, device=DEV)
p = lambda t: C.c_void_p(t.data_ptr())
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
T.truth_blend_bwd(C.c_int(W), C.c_int(H), p(o["ranges"]), p(o["point_list"]), p(o["_bg"]), p(o["_rec"]), p(o["n_contrib"]), p(dLc), p(dLd), p(acc64), st)
torch.cuda.synchronize()
names = ["mean2D.x", "mean2D.y", "conic.x", "conic.y", "conic.w", "opacity", "r", "g", "b", "invdepth"]
print("accumulators, float32 blend vs FP64 blend (max abs err / max abs):")
for q, n in enumerate(names):
    d = (acc32[:, q] - acc64[:, q]).abs()
    print(f"  {n:9s} {float(d.max() / acc64[:, q].abs().max()):.2e}   worst Gaussian {int(d.argmax())}")

def pre64(acc):
    d = lambda *s: torch.empty(s, dtype=torch.float64, device=DEV)
    g3, g2, gdc, gsh, gop, gsc, grot = d(P, 3), d(P, 2), d(P, 3), d(P, 15, 3), d(P), d(P, 3), d(P, 4)
    f32 = lambda v: float(torch.tensor(v, dtype=torch.float32))
    T.truth_preprocess_bwd(C.c_int(P), C.c_int(3), C.c_int(15), C.c_int(W), C.c_int(H), p(sc["means3D"]), p(o["radii"]), p(sc["sh"]),
                           p(o["clamped"]), p(sc["scales"]), p(sc["rotations"]), C.c_double(1.0), p(sc["viewmatrix"]), p(sc["projmatrix"]),
                           C.c_double(f32(sc["tanfovx"])), C.c_double(f32(sc["tanfovy"])), p(sc["campos"]), p(acc), C.c_double(0.0),
                           p(g3), p(g2), p(gdc), p(gsh), p(gop), p(gsc), p(grot), st)
    torch.cuda.synchronize()
    return dict(means3D=g3, scales=gsc, rotations=grot, opacities=gop.view(P, 1))

mixed = pre64(acc32)
for k in ("means3D", "opacities", "scales", "rotations"):
    t = truth[k]
    e_full = U.err_vs_truth(ours[k], t)
    e_acc = U.err_vs_truth(mixed[k].reshape(t.shape), t)
    d = (ours[k].double().reshape(t.shape) - t).abs().reshape(P, -1).max(1).values
    i = int(d.argmax())
    print(f"{k:10s} ours {e_full:.2e}   float32 accumulators + FP64 chain rule {e_acc:.2e}   worst Gaussian {i}: "
          f"scale {sc['scales'][i].tolist()} radius {int(o['radii'][i])} tiles {int(o['tiles'][i])} opacity {float(sc['opacities'][i]):.3f}")
    print("     ours", ours[k][i].flatten().tolist(), "\n     truth", t[i].flatten().tolist(), "\n     mixed", mixed[k].reshape(t.shape)[i].flatten().tolist())
    print("     acc32", acc32[i].tolist(), "\n     acc64", acc64[i].tolist())
