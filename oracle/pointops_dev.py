"""ORACLE — test infrastructure only.  Device- and dtype-generic variant of the nine `_ext` operators for running the
reference dataflow (oracle/istnet_port.py) on the GPU in FLOAT64 as ground truth at the benchmark's full shape: the
index-producing ops run the C restatement (oracle/pointops_ref.c) on an FP32 host copy of the coordinates (test inputs are
exactly representable in FP32, so the indices are the reference's), the gather-type ops are torch index arithmetic in the
tensor's own dtype on the tensor's own device.  Never imported by istnet_b200/."""
import torch

from . import pointops as po
from . import pointops_any as any_ops

gather_points = any_ops.gather_points
gather_points_grad = any_ops.gather_points_grad
group_points = any_ops.group_points
group_points_grad = any_ops.group_points_grad
three_interpolate = any_ops.three_interpolate
three_interpolate_grad = any_ops.three_interpolate_grad


def _host32(t):
    return t.detach().float().cpu().contiguous()


def furthest_point_sampling(points, nsamples):
    return po.furthest_point_sampling(_host32(points), nsamples).to(points.device)


def ball_query(new_xyz, xyz, radius, nsample):
    return po.ball_query(_host32(new_xyz), _host32(xyz), radius, nsample).to(xyz.device)


def three_nn(unknown, known):
    d2, idx = po.three_nn(_host32(unknown), _host32(known))
    d2, idx = d2.to(unknown.device), idx.to(unknown.device)
    if unknown.dtype == torch.float64:  # distances in double from the FP32-selected neighbours
        g = torch.gather(known.unsqueeze(1).expand(-1, unknown.shape[1], -1, -1), 2, idx.long().unsqueeze(-1).expand(-1, -1, -1, 3))
        d2 = (unknown.unsqueeze(2) - g).pow(2).sum(-1)
    return [d2.to(unknown.dtype), idx]
