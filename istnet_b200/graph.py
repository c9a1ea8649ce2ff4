"""CUDA-graph capture of a whole training step (forward + loss + backward).

The eager path issues ~3000 kernel launches per step through Python/ctypes and is launch-bound (≈65 ms of host time at
B=32 against ≈25 ms of GPU work); every kernel of this package is shape-static and allocation-free inside the C ABI,
so the complete step — including the TMA descriptors, which are passed by value as kernel parameters — replays as one
graph launch.  Inputs live in static device buffers that `__call__` refills.
"""
import os

import torch


class GraphedTrainStep:
    """step = GraphedTrainStep(model, loss_fn, example_batch, ...); loss = step(batch)  (gradients are left in p.grad).

    before_forward / after_backward: callables captured with the step (parallel.DataParallelStep: per-bucket gradient
    packing + NCCL all-reduce on a communication stream, Adam on the flat buckets).  before_replay / after_replay: host-side
    callables around each replay (learning-rate device scalar; re-binding `.grad` to the flat buckets).
    keep_grads=True captures with the existing `.grad` tensors (accumulating into them).

    Values a training loop changes between iterations are read from DEVICE memory by the captured kernels and refreshed here
    before every replay: the BatchNorm momentum of every nn.BatchNorm module (BNMomentumScheduler, utils/scheduler.py:277-303;
    nhwc.refresh_momentum) and, through before_replay, the learning rate."""

    def __init__(self, model, loss_fn, example, model_keys, label_keys, warmup=3, after_backward=None, keep_grads=False,
                 before_forward=None, before_replay=None, after_replay=None):
        from . import nhwc

        self.model, self.loss_fn = model, loss_fn
        self.model_keys, self.label_keys = tuple(model_keys), tuple(label_keys)
        self.static = {k: example[k].clone() for k in self.model_keys + tuple(k for k in self.label_keys if k not in self.model_keys)}
        self.before_forward, self.after_backward = before_forward, after_backward
        self.before_replay, self.after_replay = before_replay, after_replay
        self.keep_grads = keep_grads or os.environ.get("ISTNET_GRAPH_ACCUMULATE", "0") == "1"
        self._bns = [m for m in model.modules() if isinstance(m, torch.nn.modules.batchnorm._BatchNorm)]
        self._dev = next(iter(self.static.values())).device
        self._refresh = lambda: nhwc.refresh_momentum(self._bns, self._dev)
        self._refresh()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):  # warm-up off the capture stream (allocator, lazy inits, cuBLAS handles)
            for _ in range(warmup):
                self._eager()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = self._eager()
        # the gradient tensors autograd installed during the capture: graph-pool memory that every replay rewrites
        self._grads = [(p, p.grad) for p in model.parameters() if p.grad is not None]
        if self.after_replay is not None:
            self.after_replay()

    def _eager(self):
        # keep_grads: accumulate into the existing (flat-bucket) tensors, zeroed by the graph itself.  Otherwise the step
        # starts from `grad = None` (zero_grad(set_to_none=True)): autograd installs each gradient tensor it produced —
        # memory of the graph's private pool, rewritten by every replay — so the ~660 zero-fill and ~470 accumulate
        # kernels of the in-place variant disappear from the graph (0.5 ms at its head alone, tools/timeline.py).
        for p in self.model.parameters():
            if p.grad is not None:
                if self.keep_grads:
                    p.grad.zero_()
                else:
                    p.grad = None
        if self.before_forward is not None:
            self.before_forward()
        ep = self.model({k: self.static[k] for k in self.model_keys})
        ep.update({k: self.static[k] for k in self.label_keys})
        loss = self.loss_fn(ep)
        loss.backward()
        if self.after_backward is not None:
            self.after_backward()
        return loss.detach()

    def load(self, batch, non_blocking=True):
        for k, t in self.static.items():
            t.copy_(batch[k], non_blocking=non_blocking)

    def __call__(self, batch=None):
        if batch is not None:
            self.load(batch)
        self._refresh()
        if self.before_replay is not None:
            self.before_replay()
        self.graph.replay()
        if self.after_replay is not None:
            self.after_replay()  # e.g. reducer.bind_grads: `.grad` must point at the all-reduced bucket slots
        elif not self.keep_grads:
            for p, g in self._grads:  # an optimizer.zero_grad(set_to_none=True) between steps must not detach the step's output
                if p.grad is None:
                    p.grad = g
        return self.loss
