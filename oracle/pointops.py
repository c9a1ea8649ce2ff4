"""ORACLE — test infrastructure only (see oracle/__init__.py).

ctypes front-end of oracle/pointops_ref.c exposing the nine operators with the SAME names, argument
order and return conventions as the reference pybind module `pointnet2._ext`
(/root/reference/model/pointnet2/_ext_src/src/bindings.cpp:11-24), for CPU tensors.
"""
import ctypes
import os
import subprocess

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libpointops_ref.so")


def build():
    src = os.path.join(_HERE, "pointops_ref.c")
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libpointops_ref.so"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _f(t):
    assert t.dtype == torch.float32 and t.is_contiguous() and t.device.type == "cpu"
    return ctypes.c_void_p(t.data_ptr())


def _i(t):
    assert t.dtype == torch.int32 and t.is_contiguous() and t.device.type == "cpu"
    return ctypes.c_void_p(t.data_ptr())


def opt_n_threads(n):
    return lib().ref_opt_n_threads(int(n))


def furthest_point_sampling(points, nsamples):  # sampling.cpp:70-91
    b, n, _ = points.shape
    out = torch.zeros(b, nsamples, dtype=torch.int32)
    lib().ref_furthest_point_sampling(b, n, int(nsamples), _f(points), _i(out))
    return out


def gather_points(points, idx):  # sampling.cpp:20-43
    b, c, n = points.shape
    m = idx.shape[1]
    out = torch.zeros(b, c, m)
    lib().ref_gather_points(b, c, n, m, _f(points), _i(idx), _f(out))
    return out


def gather_points_grad(grad_out, idx, n):  # sampling.cpp:45-69
    b, c, m = grad_out.shape
    out = torch.zeros(b, c, n)
    lib().ref_gather_points_grad(b, c, int(n), m, _f(grad_out), _i(idx), _f(out))
    return out


def ball_query(new_xyz, xyz, radius, nsample):  # ball_query.cpp:13-37 (centroids first)
    b, m, _ = new_xyz.shape
    n = xyz.shape[1]
    out = torch.zeros(b, m, nsample, dtype=torch.int32)
    lib().ref_ball_query(b, n, m, ctypes.c_float(radius), int(nsample), _f(new_xyz), _f(xyz), _i(out))
    return out


def group_points(points, idx):  # group_points.cpp:17-40
    b, c, n = points.shape
    _, m, ns = idx.shape
    out = torch.zeros(b, c, m, ns)
    lib().ref_group_points(b, c, n, m, ns, _f(points), _i(idx), _f(out))
    return out


def group_points_grad(grad_out, idx, n):  # group_points.cpp:42-65
    b, c, m, ns = grad_out.shape
    out = torch.zeros(b, c, n)
    lib().ref_group_points_grad(b, c, int(n), m, ns, _f(grad_out), _i(idx), _f(out))
    return out


def three_nn(unknown, known):  # interpolate.cpp:19-45
    b, n, _ = unknown.shape
    m = known.shape[1]
    dist2 = torch.zeros(b, n, 3)
    idx = torch.zeros(b, n, 3, dtype=torch.int32)
    lib().ref_three_nn(b, n, m, _f(unknown), _f(known), _f(dist2), _i(idx))
    return [dist2, idx]


def three_interpolate(points, idx, weight):  # interpolate.cpp:47-75
    b, c, m = points.shape
    n = idx.shape[1]
    out = torch.zeros(b, c, n)
    lib().ref_three_interpolate(b, c, m, n, _f(points), _i(idx), _f(weight), _f(out))
    return out


def three_interpolate_grad(grad_out, idx, weight, m):  # interpolate.cpp:76-104
    b, c, n = grad_out.shape
    out = torch.zeros(b, c, m)
    lib().ref_three_interpolate_grad(b, c, n, int(m), _f(grad_out), _i(idx), _f(weight), _f(out))
    return out
