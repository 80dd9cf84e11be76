"""Golden vectors for the ADMM consensus step, produced by EXECUTING the reference's own methods
(unmodified, imported from /root/reference with tools/ref_import.py; CPU tensors):

  GaussianSplatModel.reinitialize / plus_gaussians / average_gaussians    gaussian_splat_model.py:316-340
  MasterGaussianSplatTrainer.setup_penalty_parameters                     master_gaussian_trainer.py:326-334
      .gaussian_splat_consensus / .broadcast_global_gaussian_splat        :538-555, :523-535
      .update_block_dual_variables / .set_block_penalty_parameters        :459-465, :380-393
      .compute_primal_residual / .compute_dual_residual                   :396-456
      .adapt_penalty_parameters                                           :337-377
  SlaveGaussianSplatTrainer.setup_dual_variables / update_dual_variables  slave_gaussian_trainer.py:81-121
      .set_global_gaussians / .set_penalty_parameters / .add_admm_penalties   :123-202

Scenario: K = 4 blocks over N_global Gaussians, every ownership count 1..4 present, three consensus
rounds in the order of MasterGaussianSplatTrainer.train_iteration (:697-716) with seeded "training"
perturbations of the block models in between; before rounds 2 and 3 the penalty and its autograd
gradients are evaluated per block.  The RPC layer is replaced by direct calls (a fake RRef whose
rpc_async().method(...) runs the slave's method); nothing else is altered.

    python tools/make_golden_admm.py        ->  tests/golden/admm.npz
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ref_import  # noqa: E402

GaussianSplatModel, Slave, Master, _ = ref_import.reference_classes()

NAMES = ("xyz", "fdc", "fr", "s", "q", "o")
SHAPES = ((3,), (1, 3), (15, 3), (3,), (4,), (1,))
K, N_GLOBAL = 4, 320


class _Done:
    def __init__(self, v):
        self.v = v

    def wait(self):
        return self.v


class _FakeRRef:
    """worker.rpc_async().method(args) -> future, executed in-process."""

    def __init__(self, obj):
        self.obj = obj

    def rpc_async(self):
        outer = self

        class _Proxy:
            def __getattr__(self, name):
                return lambda *a, **k: _Done(getattr(outer.obj, name)(*a, **k))
        return _Proxy()


def model_from(tensors, optimizable=False):
    m = GaussianSplatModel(max_sh_degree=3, percent_dense=0.01, device="cpu")
    m.init_from_external_properties(*[t.clone() for t in tensors], optimizable=optimizable)
    return m


def props(m):
    return [t.detach().clone() for t in m.get_all_properties()]


def main():
    g = torch.Generator(device="cpu").manual_seed(2024)
    # ownership: every Gaussian in a random non-empty subset of the 4 blocks, all counts 1..4 present
    member = torch.rand(N_GLOBAL, K, generator=g) < 0.45
    member[:8] = torch.tensor([[1, 0, 0, 0], [0, 1, 0, 0], [1, 1, 0, 0], [0, 0, 1, 1], [1, 1, 1, 0], [0, 1, 1, 1],
                               [1, 0, 1, 1], [1, 1, 1, 1]], dtype=torch.bool)
    member[member.sum(1) == 0, 0] = True
    global_indices = [torch.nonzero(member[:, k]).squeeze(1) for k in range(K)]       # ascending (argwhere, master:136)
    visibility_count = torch.bincount(torch.cat(global_indices))                      # master:168
    assert set(visibility_count.tolist()) == {1, 2, 3, 4}

    scale = (1.0, 1.0, 0.1, 1.5, 1.0, 2.0)
    z0 = [torch.randn(N_GLOBAL, *s, generator=g) * c for s, c in zip(SHAPES, scale)]

    cfg = types.SimpleNamespace(trainer=types.SimpleNamespace(admm=types.SimpleNamespace(
        alpha_xyz=1e5, alpha_fdc=1e4, alpha_fr=1e5, alpha_s=1e4, alpha_q=1e5, alpha_o=1e4,
        mu=10, tau_inc=2, tau_dec=2, over_relaxation_coeff=0.5)))

    master = object.__new__(Master)
    master.writer = None            # BaseTrainer.__del__ (trainer.py:76) looks at it
    master.config = cfg
    master.gaussians = model_from(z0)
    master.global_indices = global_indices
    master.visibility_count = visibility_count
    slaves = []
    for k in range(K):
        s = object.__new__(Slave)
        s.writer = None
        s.config = cfg
        s.scalars_to_log = {}
        s.gaussians = model_from([t[global_indices[k]] for t in z0], optimizable=True)   # get_sub_gaussians, :279-296
        slaves.append(s)
    master.remote_workers = {k: _FakeRRef(s) for k, s in enumerate(slaves)}
    master.collect_block_gaussians = lambda: {k: model_from(props(s.gaussians)) for k, s in enumerate(slaves)}
    # fuse_local_gaussians tail (master:607-616)
    master.setup_penalty_parameters(num_gaussians=N_GLOBAL)
    master.set_block_penalty_parameters()
    for s in slaves:
        s.setup_dual_variables()

    out = {"K": K, "n_global": N_GLOBAL, "visibility_count": visibility_count.numpy()}
    for k in range(K):
        out[f"global_indices_{k}"] = global_indices[k].numpy()
    for n, t in zip(NAMES, z0):
        out[f"z0_{n}"] = t.numpy()

    rho_of = lambda o: np.array([o.rho_xyz, o.rho_fdc, o.rho_fr, o.rho_s, o.rho_q, o.rho_o], dtype=np.float64)
    out["rho_init"] = rho_of(master)
    res_keys = ("xyz", "feat_dc", "feat_rest", "scale", "quaternion", "opacity")
    noise = (0.0, 0.05, 0.02)
    for rnd in range(3):
        if rnd > 0:
            # "training": seeded perturbation of every block's parameters (different per block)
            for k, s in enumerate(slaves):
                with torch.no_grad():
                    for t, c in zip(s.gaussians.get_all_properties(), scale):
                        t.add_(torch.randn(t.shape, generator=g) * (noise[rnd] * c))
            # the per-iteration penalty of the local objective and its gradient (slave:161-202)
            for k, s in enumerate(slaves):
                for t in s.gaussians.get_all_properties():
                    t.grad = None
                loss = s.add_admm_penalties(torch.zeros(()))
                loss.backward()
                out[f"r{rnd}_penalty_{k}"] = np.array(
                    [s.scalars_to_log[f"train/{n}_penalty"] for n in
                     ("xyz", "feat_dc", "feat_rest", "scaling", "quaternion", "opacity")], dtype=np.float64)
                out[f"r{rnd}_penalty_total_{k}"] = float(loss.item())
                for n, t in zip(NAMES, s.gaussians.get_all_properties()):
                    out[f"r{rnd}_pengrad_{n}_{k}"] = t.grad.numpy().copy()
        for k, s in enumerate(slaves):
            for n, t in zip(NAMES, props(s.gaussians)):
                out[f"r{rnd}_x_{n}_{k}"] = t.numpy()
        out[f"r{rnd}_rho_before"] = rho_of(master)
        # master.train_iteration steps (2)-(6), :697-716
        master.gaussian_splat_consensus()
        master.broadcast_global_gaussian_splat()
        master.update_block_dual_variables()
        primal = master.compute_primal_residual()
        dual = master.compute_dual_residual()
        master.adapt_penalty_parameters(primal_res_dict=primal, dual_res_dict=dual)
        master.set_block_penalty_parameters()
        for n, t in zip(NAMES, props(master.gaussians)):
            out[f"r{rnd}_z_{n}"] = t.numpy()
        out[f"r{rnd}_primal"] = np.array([float(primal[q]) for q in res_keys], dtype=np.float64)
        out[f"r{rnd}_dual"] = np.array([float(dual[q]) for q in res_keys], dtype=np.float64)
        out[f"r{rnd}_rho_after"] = rho_of(master)
        for k, s in enumerate(slaves):
            assert np.array_equal(rho_of(s), rho_of(master))
            for n, u in zip(NAMES, (s.u_xyz, s.u_fdc, s.u_fr, s.u_s, s.u_q, s.u_o)):
                out[f"r{rnd}_u_{n}_{k}"] = u.detach().numpy().copy()
            for n, zz in zip(NAMES, (s.global_xyz, s.global_features_dc, s.global_features_rest, s.global_scaling,
                                     s.global_quaternion, s.global_opacity)):
                assert torch.equal(zz, props(master.gaussians)[NAMES.index(n)][global_indices[k]])
    path = os.path.join(ROOT, "tests", "golden", "admm.npz")
    np.savez_compressed(path, **out)
    print("written", path, os.path.getsize(path), "bytes; stubbed:", sorted({m.split('.')[0] for m in ref_import.enable()}))
    print("rho:", out["rho_init"], "->", out["r2_rho_after"])
    print("primal r1", out["r1_primal"], "dual r1", out["r1_dual"])


if __name__ == "__main__":
    main()
