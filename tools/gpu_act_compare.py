"""dgr_apply_activations vs torch.sigmoid / torch.exp / F.normalize, bit for bit."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dogs_b200 import _lib as L
dev = "cuda:0"
P = 2_000_000
g = torch.Generator(device="cpu").manual_seed(0)
op = (torch.randn(P, 1, generator=g) * 3).to(dev)
sc = (torch.randn(P, 3, generator=g) * 1.5 - 4).to(dev)
rot = (torch.randn(P, 4, generator=g) * torch.rand(P, 1, generator=g) * 3).to(dev)
o, s, r = torch.empty_like(op), torch.empty_like(sc), torch.empty_like(rot)
L.check(L.lib().dgr_apply_activations(P, op.data_ptr(), sc.data_ptr(), rot.data_ptr(), o.data_ptr(), s.data_ptr(),
                                      r.data_ptr(), L.stream_ptr(torch.device(dev))))
torch.cuda.synchronize()
for name, mine, ref in (("sigmoid", o, torch.sigmoid(op)), ("exp", s, torch.exp(sc)),
                        ("normalize", r, torch.nn.functional.normalize(rot))):
    ne = (mine.view(torch.int32) != ref.view(torch.int32))
    print(name, "mismatching elements", int(ne.sum()), "of", mine.numel(), "max ulp",
          int((mine.view(torch.int32) - ref.view(torch.int32)).abs().max()))
# which norm does torch compute?
n_t = torch.linalg.vector_norm(rot, dim=1)
x = rot
sq = x * x
cands = {"((a+b)+c)+d": ((sq[:, 0] + sq[:, 1]) + sq[:, 2]) + sq[:, 3],
         "(a+b)+(c+d)": (sq[:, 0] + sq[:, 1]) + (sq[:, 2] + sq[:, 3]),
         "(a+c)+(b+d)": (sq[:, 0] + sq[:, 2]) + (sq[:, 1] + sq[:, 3]),
         "fma chain": torch.addcmul(torch.addcmul(torch.addcmul(sq[:, 0], x[:, 1], x[:, 1]), x[:, 2], x[:, 2]), x[:, 3], x[:, 3])}
for k, v in cands.items():
    print("norm candidate", k, "mismatch", int((torch.sqrt(v).view(torch.int32) != n_t.view(torch.int32)).sum()))
d = n_t.clamp_min(1e-12)[:, None]
print("normalize == x / norm.clamp_min:", int(((x / d).view(torch.int32) != torch.nn.functional.normalize(rot).view(torch.int32)).sum()))
