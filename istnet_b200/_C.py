"""ctypes binding of libistnet_b200.so (C ABI declared in include/istnet_b200.h).

The product path has NO fallback: if the CUDA library is missing or a call fails, a RuntimeError is raised.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libistnet_b200.so")
_lib = None

c_int, c_float, c_void_p, c_ll = ctypes.c_int, ctypes.c_float, ctypes.c_void_p, ctypes.c_longlong


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"istnet_b200: CUDA library {LIB_PATH} not built (run `python -m istnet_b200.build`); there is no CPU fallback"
            )
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.istnet_strerror.restype = ctypes.c_char_p
    return _lib


def check(status, what):
    if status != 0:
        msg = lib().istnet_strerror(int(status)).decode()
        raise RuntimeError(f"istnet_b200.{what} failed: {msg} (status {status})")


def stream_ptr():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return c_void_p(t.data_ptr())


def _require(t, dtype, name, ndim=None):
    """Mirrors CHECK_CUDA / CHECK_CONTIGUOUS / CHECK_IS_FLOAT / CHECK_IS_INT of the reference (utils.h:10-30)."""
    if not isinstance(t, torch.Tensor):
        raise RuntimeError(f"{name} must be a tensor")
    if not t.is_cuda:
        raise RuntimeError(f"{name}: CPU not supported (must be a CUDA tensor)")
    if t.dtype != dtype:
        raise RuntimeError(f"{name} must be a {'float' if dtype == torch.float32 else 'int'} tensor, got {t.dtype}")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be a contiguous tensor")
    if ndim is not None and t.dim() != ndim:
        raise RuntimeError(f"{name} must have {ndim} dimensions, got {t.dim()}")


def req_f(t, name, ndim=None):
    _require(t, torch.float32, name, ndim)


def req_i(t, name, ndim=None):
    _require(t, torch.int32, name, ndim)


class Fin(ctypes.Structure):
    """include/istnet_b200.h `istnet_fin`: in-kernel completion of a per-channel reduction by the producer's last CTA."""

    _fields_ = [
        ("kind", ctypes.c_int), ("tickets", ctypes.c_void_p), ("P", ctypes.c_longlong), ("eps", ctypes.c_float),
        ("momentum", ctypes.c_void_p), ("running_mean", ctypes.c_void_p), ("running_var", ctypes.c_void_p), ("mean", ctypes.c_void_p),
        ("invstd", ctypes.c_void_p), ("num_batches_tracked", ctypes.c_void_p), ("sum_f64", ctypes.c_void_p), ("sum_f32", ctypes.c_void_p),
        ("sum2_f32", ctypes.c_void_p),
    ]


FIN_BN_STATS, FIN_COLSUM, FIN_BN_BWD = 1, 2, 3
FIN_TICKETS, FIN_ROWS = 38, 592 + 37  # ISTNET_FIN_TICKETS / ISTNET_FIN_ROWS
_TICKET_SLOTS = 8192
_ticket_arena = {}


def tickets(dev):
    """Pointer to FIN_TICKETS zeroed uint32 counters for one reduction.  The counters are self-resetting (zero again when the
    kernel exits), so the arena is handed out round-robin: a slot is reused only thousands of launches later, and a
    captured CUDA graph keeps replaying on the slots it was captured with."""
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    a = _ticket_arena.get(key)
    if a is None:
        a = _ticket_arena[key] = [torch.zeros(_TICKET_SLOTS * FIN_TICKETS, dtype=torch.int32, device=dev), 0]
    a[1] = (a[1] + 1) % _TICKET_SLOTS
    return c_void_p(a[0].data_ptr() + 4 * FIN_TICKETS * a[1])


def _vp(t):
    return t.data_ptr() if t is not None else None


LAUNCHES = 0  # number of C-ABI calls made by this process (each enqueues >= 1 kernel); bench.py reports it


def call(name, *args):
    """Invokes `istnet_<name>` on the current device / current stream of the calling thread."""
    global LAUNCHES
    LAUNCHES += 1
    fn = getattr(lib(), "istnet_" + name)
    check(fn(*args, stream_ptr()), name)
