"""Data-parallel training plumbing: one process per GPU, parameters resident, ONE gradient all-reduce per step, Adam on
flat buckets.

Replaces the reference's only multi-GPU mechanism, `torch.nn.DataParallel` (train.py:98-99), whose per-step
replicate / scatter / gather / reduce is torch-internal, and the optimizer step of its solver (utils/solver.py:41-46,98-99).
Semantics kept: every rank computes BatchNorm statistics over its own shard (no SyncBN, as with DataParallel replicas) and the
averaged gradient equals the gradient of the full-batch mean loss (utils/solver.py:180-182).

Gradients live in flat FP32 buckets (parameters' `.grad` are views into them), ordered as backward produces them: first
everything outside the image branch (pose heads, both PointNet++ extractors: ready when a third of the backward pass is still
to run), then the image branch.  Parameters that never receive a gradient (`feats.fc`, frozen / detached sub-networks:
SURVEY.md §7 hard part 6) are discovered on the first step and left out, so their `.grad` stays None exactly as in the
reference.

Two ways to drive it:
  * eager (`zero_grad` / `backward` / `finish`): post-accumulate hooks only COUNT; every all-reduce is launched from `finish()`,
    after `loss.backward()` has returned and the autograd engine has joined the branch streams into the caller's stream
    (IST_Net.forward runs its sub-networks on side streams, so a hook firing on one of them must not start NCCL on gradients
    that another stream is still writing).
  * CUDA graph (`GraphedTrainStep(..., after_backward=trainer.graph_tail)`): the captured step starts from grad=None, packs
    the fresh gradients into the buckets with one multi-tensor copy per bucket, all-reduces each bucket on a communication
    stream as soon as its last gradient exists (the first bucket overlaps the image branch's backward), and runs Adam
    (csrc/optim.cu) on the flat buckets — all inside the graph.
"""
import torch
import torch.distributed as dist

ALIGN = 64  # elements: every parameter's slot starts 256-byte aligned inside its bucket


def _pad(n):
    return (n + ALIGN - 1) // ALIGN * ALIGN


class GradAllReducer:
    def __init__(self, module, bucket_mb=25.0, process_group=None, early=None):
        """early: optional predicate(name) -> True for parameters whose gradients are complete early in the backward pass
        (default: everything outside `rgb_cam_extractor` / `rgb_extractor`); they fill the first bucket(s)."""
        self.module = module
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.bucket_bytes = int(bucket_mb * (1 << 20))
        self.early = early or (lambda name: not (name.startswith("rgb_cam_extractor.") or name.startswith("rgb_extractor.")))
        self.buckets = None  # list of dict(flat=Tensor, params=[...], views=[...], pending=int, early=bool)
        self._hooks = []
        self._param_bucket = {}
        self._comm = None
        self._captured_reduce = set()
        self.bucket_scale = 1.0  # what the optimizer must multiply the bucket contents by to get the mean gradient
        self.tail_allreduce = True  # False: the per-bucket tail only packs (the caller all-reduces after the graph replay)

    # -- first step: plain all-reduce of whatever received a gradient, then build the buckets
    def _discover_and_build(self):
        named = [(n, p) for n, p in self.module.named_parameters() if p.requires_grad and p.grad is not None]
        for _, p in named:
            self._reduce_tensor(p.grad)
        # reverse registration order = the order backward produces them; early (non-image) parameters first
        order = [(n, p) for n, p in reversed(named) if self.early(n)] + [(n, p) for n, p in reversed(named) if not self.early(n)]
        groups, cur, cur_bytes, cur_early = [], [], 0, None
        for n, p in order:
            e = self.early(n)
            if cur and (cur_bytes >= self.bucket_bytes or e != cur_early):
                groups.append((cur, cur_early))
                cur, cur_bytes = [], 0
            cur.append(p)
            cur_early = e
            cur_bytes += p.numel() * p.element_size()
        if cur:
            groups.append((cur, cur_early))
        self.buckets = []
        for bi, (params, e) in enumerate(groups):
            flat = torch.zeros(sum(_pad(p.numel()) for p in params), dtype=params[0].dtype, device=params[0].device)
            views, off = [], 0
            for p in params:
                v = flat[off : off + p.numel()].view_as(p)
                v.copy_(p.grad)
                p.grad = v
                views.append(v)
                off += _pad(p.numel())
                self._param_bucket[p] = bi
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))
            self.buckets.append({"flat": flat, "params": params, "views": views, "pending": len(params), "early": e})

    def _reduce_tensor(self, t):
        if self.world == 1:
            return
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        t.div_(self.world)

    def _on_grad(self, p):
        b = self.buckets[self._param_bucket[p]]
        b["pending"] -= 1
        if b["pending"] == 0 and self._capture_tail is not None:
            self._capture_tail(b)

    _capture_tail = None

    def zero_grad(self):
        if self.buckets is None:
            self.module.zero_grad(set_to_none=True)
            return
        for b in self.buckets:
            b["flat"].zero_()
            b["pending"] = len(b["params"])
        self.bind_grads()  # an optimizer.zero_grad(set_to_none=True) in between must not detach the parameters from the buckets

    def _pack(self, b):
        """Copies gradients that autograd left outside their bucket slot (after `zero_grad(set_to_none=True)`, or a step that
        started from grad=None) into the slot and re-binds `.grad` to it."""
        src, dst = [], []
        for p, v in zip(b["params"], b["views"]):
            if p.grad is not None and p.grad.data_ptr() != v.data_ptr():
                src.append(p.grad)
                dst.append(v)
        if src:
            torch._foreach_copy_(dst, src)
        for p, v in zip(b["params"], b["views"]):
            if p.grad is not None and p.grad.data_ptr() != v.data_ptr():
                p.grad = v

    def finish(self):
        """Call after loss.backward(): packs stray gradients, all-reduces every bucket (all launched back to back, after the
        autograd engine has synchronised the branch streams with the caller's stream) and averages."""
        if self.buckets is None:
            self._discover_and_build()
            return
        hs = []
        for b in self.buckets:
            self._pack(b)
            if self.world > 1:
                hs.append(dist.all_reduce(b["flat"], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        for h, b in zip(hs, self.buckets):
            h.wait()
            b["flat"].div_(self.world)
        for b in self.buckets:
            b["pending"] = len(b["params"])
        self.bucket_scale = 1.0

    def reduce_all(self, average=True):
        """All-reduce every bucket now (after a CUDA-graph replay of forward+backward whose tail packed the buckets)."""
        if self.buckets is None:
            self._discover_and_build()
            return
        if self.world == 1:
            return
        hs = [dist.all_reduce(b["flat"], op=dist.ReduceOp.SUM, group=self.group, async_op=True) for b in self.buckets]
        for h, b in zip(hs, self.buckets):
            h.wait()
            if average:
                b["flat"].div_(self.world)
        self.bucket_scale = 1.0 if average else 1.0 / self.world

    # -- CUDA-graph path: the captured step starts from grad=None, so autograd leaves every gradient in a tensor of the
    # graph's private pool (no zero-fill, no accumulate kernels); one multi-tensor copy per bucket packs them into the flat
    # buckets
    def gather_grads(self, only=None):
        """Copies every parameter's current `.grad` into its bucket slot (call inside the capture, after backward)."""
        if self.buckets is None:
            return
        for b in self.buckets if only is None else only:
            src, dst = [], []
            for p, v in zip(b["params"], b["views"]):
                if p.grad is not None and p.grad.data_ptr() != v.data_ptr():
                    src.append(p.grad)
                    dst.append(v)
            if src:
                torch._foreach_copy_(dst, src)

    def bind_grads(self):
        """Points every `.grad` at its bucket slot (the optimizer then reads the all-reduced values)."""
        if self.buckets is None:
            return
        for b in self.buckets:
            for p, v in zip(b["params"], b["views"]):
                if p.grad is None or p.grad.data_ptr() != v.data_ptr():
                    p.grad = v

    def remove_hooks(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []

    def num_buckets(self):
        return 0 if self.buckets is None else len(self.buckets)

    def grad_bytes(self):
        return 0 if self.buckets is None else sum(b["flat"].numel() * 4 for b in self.buckets)

    # -- all-reduce inside the capture, on a communication stream, per bucket as it completes
    def comm_stream(self, dev):
        if self._comm is None:
            self._comm = torch.cuda.Stream(dev)
        return self._comm

    def begin_capture_overlap(self, streams):
        """Arms the per-bucket tail for a step that is being captured (or run eagerly) from grad=None: when the last gradient of
        a bucket has been produced, the bucket is packed and all-reduced on the communication stream, which first waits for
        every stream in `streams` (the branch streams of IST_Net.forward; gradients of one bucket come from several of them)."""
        self._captured_reduce = set()
        for b in self.buckets:
            b["pending"] = len(b["params"])

        def _active(st):  # under capture only streams that already belong to the capture may be waited on
            if not torch.cuda.is_current_stream_capturing():
                return True
            with torch.cuda.stream(st):
                return torch.cuda.is_current_stream_capturing()

        def tail(b):
            dev = b["flat"].device
            if dev.type != "cuda":  # host tensors (gloo tests): no streams to order
                self.gather_grads(only=[b])
                if self.world > 1 and self.tail_allreduce:
                    dist.all_reduce(b["flat"], op=dist.ReduceOp.SUM, group=self.group)
                self._captured_reduce.add(id(b))
                return
            comm = self.comm_stream(dev)
            comm.wait_stream(torch.cuda.current_stream(dev))
            for st in streams:
                if _active(st):
                    comm.wait_stream(st)
            with torch.cuda.stream(comm):
                self.gather_grads(only=[b])
                if self.world > 1 and self.tail_allreduce:
                    dist.all_reduce(b["flat"], op=dist.ReduceOp.SUM, group=self.group)
            self._captured_reduce.add(id(b))

        self._capture_tail = tail

    def end_capture_overlap(self):
        """Buckets whose hooks did not all fire are packed / reduced now; the current stream then waits for the communication
        stream.  After this the flat buckets hold the gradient SUM over ranks (Adam's grad_scale = 1/world averages)."""
        tail, self._capture_tail = self._capture_tail, None
        for b in self.buckets:
            if id(b) not in self._captured_reduce:
                tail(b)
        dev = self.buckets[0]["flat"].device
        if dev.type == "cuda":
            torch.cuda.current_stream(dev).wait_stream(self._comm)
        self.bucket_scale = 1.0 / self.world


class FlatAdam(torch.optim.Optimizer):
    """torch.optim.Adam (utils/solver.py:41-44) on the reducer's flat buckets: parameters are re-pointed into flat buffers
    with the buckets' layout and one kernel per bucket (csrc/optim.cu) updates them from the flat gradients.  It is a
    torch Optimizer, so CyclicLR / checkpointing code written against `optimizer.param_groups` keeps working; `lr` is
    mirrored into a device scalar before every step, which lets `step()` be captured in a CUDA graph (`sync_lr()` is then
    called before each replay)."""

    def __init__(self, reducer, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        assert reducer.buckets is not None, "run one step first: the buckets are built from the parameters that receive gradients"
        params = [p for b in reducer.buckets for p in b["params"]]
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.reducer = reducer
        dev = params[0].device
        self.flat = []
        for b in reducer.buckets:
            fp = torch.zeros_like(b["flat"])
            off = 0
            for p in b["params"]:
                v = fp[off : off + p.numel()].view_as(p)
                v.copy_(p.data)
                p.data = v
                off += _pad(p.numel())
            self.flat.append({"p": fp, "g": b["flat"], "m": torch.zeros_like(fp), "v": torch.zeros_like(fp)})
        self.lr_dev = torch.zeros(1, dtype=torch.float32, device=dev)
        # pinned ring: a training loop that runs ahead of the device rewrites lr every iteration (CyclicLR) while earlier
        # asynchronous copies may still be pending; a slot is reused only after the copy that read it has completed
        self.lr_host = torch.zeros(self._LR_RING, dtype=torch.float32).pin_memory() if dev.type == "cuda" else torch.zeros(self._LR_RING)
        self._lr_events = [None] * self._LR_RING
        self._lr_slot = 0
        self._lr_mirror = None
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        self.sync_lr()

    _LR_RING = 16

    def sync_lr(self):
        lr = float(self.param_groups[0]["lr"])
        if lr != self._lr_mirror:
            self._lr_mirror = lr
            i = self._lr_slot = (self._lr_slot + 1) % self._LR_RING
            if self._lr_events[i] is not None:
                self._lr_events[i].synchronize()
            self.lr_host[i] = lr
            self.lr_dev.copy_(self.lr_host[i : i + 1], non_blocking=True)
            if self.lr_dev.is_cuda:
                ev = self._lr_events[i] or torch.cuda.Event()
                ev.record()
                self._lr_events[i] = ev

    def zero_grad(self, set_to_none=True):
        # gradients are views of the flat buckets that every step overwrites (graph path) or that the reducer zeroes (eager
        # path): detaching them (set_to_none) would silently disconnect the optimizer from the all-reduced values
        if self.reducer._capture_tail is None and not torch.cuda.is_current_stream_capturing():
            self.reducer.zero_grad()

    @torch.no_grad()
    def step(self, closure=None):
        from . import _C
        from ctypes import c_double

        from ._C import c_float, c_ll, ptr

        g0 = self.param_groups[0]
        if not torch.cuda.is_current_stream_capturing():
            self.sync_lr()
        b1, b2 = g0["betas"]
        scale = self.reducer.bucket_scale
        for f in self.flat:
            _C.call("adam_flat", ptr(f["p"]), ptr(f["g"]), ptr(f["m"]), ptr(f["v"]), c_ll(f["p"].numel()), ptr(self.lr_dev), c_double(b1),
                    c_double(b2), c_float(g0["eps"]), c_float(g0["weight_decay"]), c_float(scale), ptr(self.step_dev))
        _C.call("adam_tick", ptr(self.step_dev))
        from . import nhwc

        nhwc.invalidate_weights(self.step_dev.device)  # parameters changed behind autograd's version counters


def broadcast_module(module, src=0, process_group=None):
    """Rank `src` parameters + buffers to every rank (start-of-training sync; DataParallel does this every step)."""
    if not dist.is_initialized() or dist.get_world_size(process_group) == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src, group=process_group)


class DataParallelStep:
    """Callbacks for GraphedTrainStep that put the rest of the training step into the captured graph: gradient buckets packed
    and all-reduced on the communication stream as they complete (`before` arms the hooks, `after` joins), then Adam on the flat
    buckets.  nccl_in_graph=False keeps NCCL out of the capture: the graph then ends after the bucket packing and the
    all-reduce + Adam kernels are issued after each replay (`after_replay`)."""

    def __init__(self, reducer, optimizer=None, nccl_in_graph=True):
        self.reducer, self.optimizer = reducer, optimizer
        self.in_graph = nccl_in_graph or reducer.world == 1

    def _streams(self):
        from . import model as M
        from . import nhwc

        out = []
        for pool in M._Branches._pool.values():
            out += list(pool)
        out += list(nhwc._SIDE.values())
        return out

    def before(self):
        self.reducer.tail_allreduce = self.in_graph
        self.reducer.begin_capture_overlap(self._streams())

    def after(self):
        self.reducer.end_capture_overlap()
        if self.in_graph and self.optimizer is not None:
            self.optimizer.step()

    def before_replay(self):
        if self.optimizer is not None:
            self.optimizer.sync_lr()

    def after_replay(self):
        if not self.in_graph:
            self.reducer.reduce_all(average=False)
            if self.optimizer is not None:
                self.optimizer.step()
        self.reducer.bind_grads()
