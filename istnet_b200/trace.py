"""Phase timeline of a (graph-replayed, multi-stream) step: `mark(name)` enqueues a one-thread kernel on the current
stream that stores the device's %globaltimer; inside a captured CUDA graph the markers are ordinary nodes, so after a
replay `report()` shows when each phase of each concurrent branch really started.  Off unless `enable()` was called
(tools/timeline.py); never active in bench.py's timed region."""
import ctypes

import torch

from . import _C

_buf = None
_names = []


def enable(slots=1024):
    global _buf
    _buf = torch.zeros(slots, dtype=torch.int64, device="cuda")
    _names.clear()


def reset():
    _names.clear()


def mark(name):
    if _buf is None or len(_names) >= _buf.numel():
        return
    slot = len(_names)
    _names.append((name, torch.cuda.current_stream().cuda_stream))
    _C.call("marker", _C.ptr(_buf), ctypes.c_int(slot))
    _C.LAUNCHES -= 1  # instrumentation, not part of the step


def report():
    torch.cuda.synchronize()
    t = _buf[: len(_names)].cpu().tolist()
    t0 = min(t)
    streams = {}
    rows = []
    for (name, st), v in zip(_names, t):
        sid = streams.setdefault(st, len(streams))
        rows.append(((v - t0) / 1e6, sid, name))
    return sorted(rows)


class _Tag(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, name):
        ctx.name = name
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        mark(ctx.name)
        return g, None


def tag_bwd(x, name):
    """Identity whose backward drops a marker: placed on a phase's output it stamps the start of that phase's backward."""
    if _buf is None or not isinstance(x, torch.Tensor) or not x.requires_grad:
        return x
    return _Tag.apply(x, name)


def phase(name, fn):
    """Runs fn() between two forward markers and tags its tensor outputs for the backward timeline."""
    if _buf is None:
        return fn()
    mark(name + " fwd>")
    out = fn()
    if isinstance(out, (tuple, list)):
        out = type(out)(tag_bwd(o, name + " bwd>") for o in out)
    else:
        out = tag_bwd(out, name + " bwd>")
    mark(name + " fwd<")
    return out
