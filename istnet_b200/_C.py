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


LAUNCHES = 0  # number of C-ABI calls made by this process (each enqueues >= 1 kernel); bench.py reports it


def call(name, *args):
    """Invokes `istnet_<name>` on the current device / current stream of the calling thread."""
    global LAUNCHES
    LAUNCHES += 1
    fn = getattr(lib(), "istnet_" + name)
    check(fn(*args, stream_ptr()), name)
