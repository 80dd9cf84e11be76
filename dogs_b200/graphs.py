"""A training step (render -> loss -> backward [-> optimizer]) replayed as ONE CUDA graph.

Small scenes are host-bound: at 100 k Gaussians the device needs 0.47 ms per step and Python + ~45 launches take
0.65 ms to issue it.  The rasterizer forward normally reads the instance count back at its end, which cannot be
captured; in a capture it runs in the library's deferred mode instead (DgrParams::deferred_check): capacity and
depth-key bits learnt from the eager warm-up calls are assumed, the measured values are copied to pinned memory
inside the graph, and this class checks them after every replay.  If an assumption was violated (the scene grew)
the step is repeated eagerly — so the caller never sees an invalid result — and the graph is captured again.

    step = GraphedStep(fn)       # fn(): set grads to None, render, loss, backward (static input tensors);
    loss = step()                #       returns a tensor (or tuple of tensors) that lives in the graph's pool

Inputs that change between steps (camera matrices, target image) must be updated IN PLACE (tensor.copy_), as with
any CUDA graph.  Leaf .grad tensors are allocated by the captured backward and stay at fixed addresses.
fn must not modify the parameters (keep optimizer.step() outside, after the call has returned): a replay whose
assumptions failed produces invalid gradients before it is detected and repeated."""
import torch

from . import diff_gaussian_rasterization as _dgr


class GraphedStep:
    def __init__(self, fn, warmup: int = 3, slots: int = 16):
        self.fn = fn
        self.slots = slots
        self.recaptures = 0
        # Eager calls run on a side stream, as torch's whole-network capture recipe asks: autograd's AccumulateGrad
        # nodes must not be tied to the default stream, or the captured backward would synchronise with it.
        self._side = torch.cuda.Stream()
        for _ in range(max(int(warmup), 1)):
            self._eager()                         # learns capacity / key bits, warms the allocator
        self._capture()

    def _eager(self):
        cur = torch.cuda.current_stream()
        self._side.wait_stream(cur)
        with torch.cuda.stream(self._side):
            out = self.fn()
        cur.wait_stream(self._side)
        torch.cuda.synchronize()
        return out

    def _capture(self):
        torch.cuda.synchronize()
        self.checks = _dgr.begin_deferred_checks(self.slots)
        try:
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.out = self.fn()
        finally:
            _dgr.end_deferred_checks()

    def __call__(self):
        for _ in range(3):
            self.graph.replay()
            torch.cuda.current_stream().synchronize()
            if self.checks.ok():
                return self.out
            # The captured capacity / key range no longer holds: this replay's outputs are invalid.  Run the step
            # eagerly (exact; learns the new capacity), capture again and replay the new graph.
            del self.out
            self._eager()
            self.recaptures += 1
            self._capture()
        raise RuntimeError("GraphedStep: the instance count keeps outgrowing the captured capacity")
