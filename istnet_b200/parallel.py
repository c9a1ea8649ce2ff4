"""Data-parallel training plumbing: one process per GPU, parameters resident, ONE gradient all-reduce per step.

Replaces the reference's only multi-GPU mechanism, `torch.nn.DataParallel` (train.py:98-99), whose per-step
replicate / scatter / gather / reduce is torch-internal.  Semantics kept: every rank computes BatchNorm statistics
over its own shard (no SyncBN, as with DataParallel replicas) and the averaged gradient equals the gradient of the
full-batch mean loss (utils/solver.py:180-182).

Gradients live in flat FP32 buckets (parameters' `.grad` are views into them), filled in reverse registration
order — the order backward produces them — and each bucket's all-reduce is launched from a post-accumulate hook as
soon as its last gradient lands, so NCCL traffic over NVLink overlaps the rest of backward.  Parameters that never
receive a gradient (`feats.fc`, frozen / detached sub-networks: SURVEY.md §7 hard part 6) are discovered on the first
step and left out, so their `.grad` stays None exactly as in the reference.
"""
import torch
import torch.distributed as dist


class GradAllReducer:
    def __init__(self, module, bucket_mb=25.0, process_group=None):
        self.module = module
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.bucket_bytes = int(bucket_mb * (1 << 20))
        self.buckets = None  # list of dict(flat=Tensor, params=[...], pending=int)
        self._handles = []
        self._hooks = []
        self._param_bucket = {}

    # -- first step: plain all-reduce of whatever received a gradient, then build the buckets
    def _discover_and_build(self):
        used = [p for p in self.module.parameters() if p.requires_grad and p.grad is not None]
        for p in used:
            self._reduce_tensor(p.grad)
        order = list(reversed(used))
        self.buckets, cur, cur_bytes = [], [], 0
        for p in order:
            cur.append(p)
            cur_bytes += p.numel() * p.element_size()
            if cur_bytes >= self.bucket_bytes:
                self.buckets.append(cur)
                cur, cur_bytes = [], 0
        if cur:
            self.buckets.append(cur)
        built = []
        for bi, params in enumerate(self.buckets):
            flat = torch.zeros(sum(p.numel() for p in params), dtype=params[0].dtype, device=params[0].device)
            off = 0
            for p in params:
                view = flat[off : off + p.numel()].view_as(p)
                view.copy_(p.grad)
                p.grad = view
                off += p.numel()
                self._param_bucket[p] = bi
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))
            built.append({"flat": flat, "params": params, "pending": len(params)})
        self.buckets = built

    def _reduce_tensor(self, t, async_op=False):
        if self.world == 1:
            return None
        h = dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group, async_op=async_op)
        if not async_op:
            t.div_(self.world)
        return h

    def _on_grad(self, p):
        b = self.buckets[self._param_bucket[p]]
        b["pending"] -= 1
        if b["pending"] == 0 and self.world > 1:
            self._handles.append((dist.all_reduce(b["flat"], op=dist.ReduceOp.SUM, group=self.group, async_op=True), b))

    def zero_grad(self):
        if self.buckets is None:
            self.module.zero_grad(set_to_none=True)
            return
        for b in self.buckets:
            b["flat"].zero_()
            b["pending"] = len(b["params"])

    def finish(self):
        """Call after loss.backward(): waits for / launches the remaining all-reduces and averages."""
        if self.buckets is None:
            self._discover_and_build()
            return
        launched = {id(b) for _, b in self._handles}
        for b in self.buckets:  # a bucket whose parameters did not all fire this step
            if id(b) not in launched and self.world > 1:
                self._handles.append((dist.all_reduce(b["flat"], op=dist.ReduceOp.SUM, group=self.group, async_op=True), b))
        for h, b in self._handles:
            h.wait()
            b["flat"].div_(self.world)
        self._handles = []

    def reduce_all(self):
        """All-reduce every bucket now (used after a CUDA-graph replay of forward+backward, where the hooks do not run):
        the bucket all-reduces are enqueued back to back and overlap each other on the NCCL stream."""
        if self.buckets is None:
            self._discover_and_build()
            return
        if self.world == 1:
            return
        hs = [dist.all_reduce(b["flat"], op=dist.ReduceOp.SUM, group=self.group, async_op=True) for b in self.buckets]
        for h, b in zip(hs, self.buckets):
            h.wait()
            b["flat"].div_(self.world)

    # -- CUDA-graph path: the captured step starts from grad=None, so autograd leaves every gradient in a tensor of the
    # graph's private pool (no zero-fill, no accumulate kernels); ONE multi-tensor copy, captured as the graph's last node,
    # packs them into the flat buckets that NCCL reduces after the replay
    def _views(self):
        for b in self.buckets:
            off = 0
            for p in b["params"]:
                yield p, b["flat"][off : off + p.numel()].view_as(p)
                off += p.numel()

    def gather_grads(self):
        """Copies every parameter's current `.grad` into its bucket slot (call inside the capture, after backward)."""
        if self.buckets is None:
            return
        src, dst = [], []
        for p, v in self._views():
            if p.grad is not None and p.grad.data_ptr() != v.data_ptr():
                src.append(p.grad)
                dst.append(v)
        if src:
            torch._foreach_copy_(dst, src)

    def bind_grads(self):
        """Points every `.grad` at its bucket slot (after the capture: the optimizer then reads the all-reduced values)."""
        if self.buckets is None:
            return
        for p, v in self._views():
            p.grad = v

    def remove_hooks(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []

    def num_buckets(self):
        return 0 if self.buckets is None else len(self.buckets)

    def grad_bytes(self):
        return 0 if self.buckets is None else sum(b["flat"].numel() * 4 for b in self.buckets)


def broadcast_module(module, src=0, process_group=None):
    """Rank `src` parameters + buffers to every rank (start-of-training sync; DataParallel does this every step)."""
    if not dist.is_initialized() or dist.get_world_size(process_group) == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src, group=process_group)
