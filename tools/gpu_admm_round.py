"""Device time of the two kernels of a consensus round, per block, on the C4 layout (8 blocks on one GPU):
    python tools/gpu_admm_round.py [C4] [per_block]"""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dogs_b200 import admm as A, scenes
cfg = sys.argv[1] if len(sys.argv) > 1 else "C4"
per_block = int(sys.argv[2]) if len(sys.argv) > 2 else None
dev = "cuda:0"
xyz, gidx, _ = scenes.admm_layout(cfg, per_block=per_block)
n_global = xyz.shape[0]
g = torch.Generator(device="cpu").manual_seed(0)
params = {b: [torch.randn(gidx[b].numel(), sz, generator=g).to(dev) for sz in A.GROUP_SIZES] for b in range(8)}
topo = A.build_topology({b: gidx[b].to(dev) for b in range(8)}, n_global)
grp = A.ConsensusGroup(params, topo)
ops = grp.ops
def timed(fn, reps=5):
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]
rep = {"cfg": cfg, "n_global": n_global, "blocks": []}
for b in range(8):
    blk = grp.blocks[b]; p = blk.plan
    out = grp.sendbuf[p.send_start:p.send_start + p.send_rows]
    t_pack = timed(lambda: ops.pack_rows([t.detach() for t in blk.x], p, out))
    t_upd = timed(lambda: ops.consensus_update([t.detach() for t in blk.x], blk.u, blk.z, blk.zmu, p, [grp.sendbuf, grp.recvbuf], grp.relax, grp.scratch, blk.residuals))
    n, S, R = p.n_local, p.S, p.send_rows
    pack_bytes = R * 236 * 2 + R * 4
    upd_bytes = n * 236 * 4 + S * 236 * 2 + R * 236 + n * 4 + S * (4 * p.max_addends + 1)     # x, z r/w, zmu w; u r/w (shared); addend rows
    rep["blocks"].append({"block": b, "n": n, "shared": S, "send_rows": R, "max_addends": p.max_addends, "pack_ms": t_pack, "update_ms": t_upd,
                          "pack_gbs": pack_bytes / t_pack / 1e6, "update_gbs": upd_bytes / t_upd / 1e6})
    print(rep["blocks"][-1])
t_round = timed(lambda: grp.consensus(adapt=False), 3)
rep["round_ms"] = t_round
print("round", t_round)
json.dump(rep, open(os.path.join(ROOT, "gpurun_out", "admm_round.json"), "w"), indent=1)
