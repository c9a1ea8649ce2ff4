"""The nine operators of the reference pybind module `pointnet2._ext`, on the B200 kernels.

Same names, argument order, dtypes, shapes, return conventions and error type (RuntimeError) as
/root/reference/model/pointnet2/_ext_src/src/bindings.cpp:11-24, so the reference's
`pointnet2_utils.py` runs unchanged with `sys.modules['pointnet2._ext'] = istnet_b200.ext`
(see compat/pointnet2/_ext.py and INTEGRATION.md).  Outputs are allocated here with torch (caller owns them),
the work is enqueued on the current CUDA stream of the calling thread.
"""
import torch

from . import _C
from ._C import c_float, c_int, ptr, req_f, req_i


def _same_device(*ts):
    d = ts[0].device
    for t in ts[1:]:
        if t.device != d:
            raise RuntimeError("all tensors must be on the same CUDA device")


def furthest_point_sampling(points, nsamples):
    """sampling.cpp:70-91 — points f32[B,N,3] -> int32[B,nsamples]"""
    req_f(points, "points", 3)
    if points.shape[2] != 3:
        raise RuntimeError("points must be (B, N, 3)")
    b, n, _ = points.shape
    out = torch.zeros(b, int(nsamples), dtype=torch.int32, device=points.device)
    with torch.cuda.device_of(points):
        _C.call("furthest_point_sampling", c_int(b), c_int(n), c_int(int(nsamples)), ptr(points), ptr(out))
    return out


def gather_points(points, idx):
    """sampling.cpp:20-43 — points f32[B,C,N], idx i32[B,m] -> f32[B,C,m]"""
    req_f(points, "points", 3)
    req_i(idx, "idx", 2)
    _same_device(points, idx)
    b, c, n = points.shape
    m = idx.shape[1]
    out = torch.empty(b, c, m, dtype=torch.float32, device=points.device)
    with torch.cuda.device_of(points):
        _C.call("gather_points", c_int(b), c_int(c), c_int(n), c_int(m), ptr(points), ptr(idx), ptr(out))
    return out


def gather_points_grad(grad_out, idx, n):
    """sampling.cpp:45-69 — grad f32[B,C,m], idx i32[B,m], n -> f32[B,C,n]"""
    req_f(grad_out, "grad_out", 3)
    req_i(idx, "idx", 2)
    _same_device(grad_out, idx)
    b, c, m = grad_out.shape
    out = torch.zeros(b, c, int(n), dtype=torch.float32, device=grad_out.device)
    with torch.cuda.device_of(grad_out):
        _C.call("gather_points_grad", c_int(b), c_int(c), c_int(int(n)), c_int(m), ptr(grad_out), ptr(idx), ptr(out))
    return out


def ball_query(new_xyz, xyz, radius, nsample):
    """ball_query.cpp:13-37 — centroids FIRST: new_xyz f32[B,m,3], xyz f32[B,N,3] -> int32[B,m,nsample]"""
    req_f(new_xyz, "new_xyz", 3)
    req_f(xyz, "xyz", 3)
    _same_device(new_xyz, xyz)
    b, m, _ = new_xyz.shape
    n = xyz.shape[1]
    out = torch.empty(b, m, int(nsample), dtype=torch.int32, device=xyz.device)
    with torch.cuda.device_of(xyz):
        _C.call("ball_query", c_int(b), c_int(n), c_int(m), c_float(float(radius)), c_int(int(nsample)), ptr(new_xyz), ptr(xyz), ptr(out))
    return out


def group_points(points, idx):
    """group_points.cpp:17-40 — points f32[B,C,N], idx i32[B,m,ns] -> f32[B,C,m,ns]"""
    req_f(points, "points", 3)
    req_i(idx, "idx", 3)
    _same_device(points, idx)
    b, c, n = points.shape
    _, m, ns = idx.shape
    out = torch.empty(b, c, m, ns, dtype=torch.float32, device=points.device)
    with torch.cuda.device_of(points):
        _C.call("group_points", c_int(b), c_int(c), c_int(n), c_int(m), c_int(ns), ptr(points), ptr(idx), ptr(out))
    return out


def group_points_grad(grad_out, idx, n):
    """group_points.cpp:42-65 — grad f32[B,C,m,ns], idx, N -> f32[B,C,N]"""
    req_f(grad_out, "grad_out", 4)
    req_i(idx, "idx", 3)
    _same_device(grad_out, idx)
    b, c, m, ns = grad_out.shape
    out = torch.zeros(b, c, int(n), dtype=torch.float32, device=grad_out.device)
    with torch.cuda.device_of(grad_out):
        _C.call("group_points_grad", c_int(b), c_int(c), c_int(int(n)), c_int(m), c_int(ns), ptr(grad_out), ptr(idx), ptr(out))
    return out


def three_nn(unknown, known):
    """interpolate.cpp:19-45 — unknown f32[B,n,3], known f32[B,m,3] -> [dist2 f32[B,n,3], idx i32[B,n,3]]"""
    req_f(unknown, "unknowns", 3)
    req_f(known, "knows", 3)
    _same_device(unknown, known)
    b, n, _ = unknown.shape
    m = known.shape[1]
    dist2 = torch.empty(b, n, 3, dtype=torch.float32, device=unknown.device)
    idx = torch.empty(b, n, 3, dtype=torch.int32, device=unknown.device)
    with torch.cuda.device_of(unknown):
        _C.call("three_nn", c_int(b), c_int(n), c_int(m), ptr(unknown), ptr(known), ptr(dist2), ptr(idx))
    return [dist2, idx]


def three_interpolate(points, idx, weight):
    """interpolate.cpp:47-75 — points f32[B,c,m], idx i32[B,n,3], weight f32[B,n,3] -> f32[B,c,n]"""
    req_f(points, "points", 3)
    req_i(idx, "idx", 3)
    req_f(weight, "weight", 3)
    _same_device(points, idx, weight)
    b, c, m = points.shape
    n = idx.shape[1]
    out = torch.empty(b, c, n, dtype=torch.float32, device=points.device)
    with torch.cuda.device_of(points):
        _C.call("three_interpolate", c_int(b), c_int(c), c_int(m), c_int(n), ptr(points), ptr(idx), ptr(weight), ptr(out))
    return out


def three_interpolate_grad(grad_out, idx, weight, m):
    """interpolate.cpp:76-104 — grad f32[B,c,n], idx, weight, m -> f32[B,c,m]"""
    req_f(grad_out, "grad_out", 3)
    req_i(idx, "idx", 3)
    req_f(weight, "weight", 3)
    _same_device(grad_out, idx, weight)
    b, c, n = grad_out.shape
    out = torch.zeros(b, c, int(m), dtype=torch.float32, device=grad_out.device)
    with torch.cuda.device_of(grad_out):
        _C.call("three_interpolate_grad", c_int(b), c_int(c), c_int(n), c_int(int(m)), ptr(grad_out), ptr(idx), ptr(weight), ptr(out))
    return out


def fps_chain(xyz, npoints):
    """FPS + centroid gather for all SA levels of one extractor in one launch (see include/istnet_b200.h §2).

    xyz f32[B,N,3] -> ([idx_l int32[B,npoint_l]], [new_xyz_l f32[B,npoint_l,3]])
    """
    import ctypes

    req_f(xyz, "xyz", 3)
    b, n, _ = xyz.shape
    L = len(npoints)
    idxs = [torch.empty(b, int(m), dtype=torch.int32, device=xyz.device) for m in npoints]
    xyzs = [torch.empty(b, int(m), 3, dtype=torch.float32, device=xyz.device) for m in npoints]
    np_arr = (ctypes.c_int * L)(*[int(m) for m in npoints])
    idx_arr = (ctypes.c_void_p * L)(*[t.data_ptr() for t in idxs])
    xyz_arr = (ctypes.c_void_p * L)(*[t.data_ptr() for t in xyzs])
    with torch.cuda.device_of(xyz):
        _C.call("fps_chain", c_int(b), c_int(n), c_int(L), np_arr, ptr(xyz), idx_arr, xyz_arr)
    return idxs, xyzs
