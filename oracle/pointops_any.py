"""ORACLE — test infrastructure only.  dtype-generic variant of the nine `_ext` operators: index-producing ops run the
C restatement on an FP32 cast of the coordinates (test inputs are exactly representable), the gather-type ops are
torch index arithmetic in the tensor's own dtype.  Lets the reference modules run in FLOAT64 to provide a ground
truth against which both the reference's FP32 result and the CUDA path are measured (tests/gen_golden.py)."""
import torch

from . import pointops as po


def furthest_point_sampling(points, nsamples):
    return po.furthest_point_sampling(points.float().contiguous(), nsamples)


def ball_query(new_xyz, xyz, radius, nsample):
    return po.ball_query(new_xyz.float().contiguous(), xyz.float().contiguous(), radius, nsample)


def three_nn(unknown, known):
    d2, idx = po.three_nn(unknown.float().contiguous(), known.float().contiguous())
    if unknown.dtype == torch.float64:  # recompute the distances in double from the FP32-selected neighbours
        g = torch.gather(known.unsqueeze(1).expand(-1, unknown.shape[1], -1, -1), 2, idx.long().unsqueeze(-1).expand(-1, -1, -1, 3))
        d2 = (unknown.unsqueeze(2) - g).pow(2).sum(-1)
    return [d2, idx]


def gather_points(points, idx):
    return torch.gather(points, 2, idx.long().unsqueeze(1).expand(-1, points.shape[1], -1))


def gather_points_grad(grad_out, idx, n):
    out = grad_out.new_zeros(grad_out.shape[0], grad_out.shape[1], n)
    return out.scatter_add_(2, idx.long().unsqueeze(1).expand(-1, grad_out.shape[1], -1), grad_out)


def group_points(points, idx):
    b, c, n = points.shape
    _, m, ns = idx.shape
    return torch.gather(points, 2, idx.long().reshape(b, 1, m * ns).expand(-1, c, -1)).view(b, c, m, ns)


def group_points_grad(grad_out, idx, n):
    b, c, m, ns = grad_out.shape
    out = grad_out.new_zeros(b, c, n)
    return out.scatter_add_(2, idx.long().reshape(b, 1, m * ns).expand(-1, c, -1), grad_out.reshape(b, c, m * ns))


def three_interpolate(points, idx, weight):
    c = points.shape[1]
    out = 0
    for q in range(3):
        out = out + torch.gather(points, 2, idx[..., q].long().unsqueeze(1).expand(-1, c, -1)) * weight[..., q].unsqueeze(1)
    return out


def three_interpolate_grad(grad_out, idx, weight, m):
    b, c, n = grad_out.shape
    out = grad_out.new_zeros(b, c, m)
    for q in range(3):
        out.scatter_add_(2, idx[..., q].long().unsqueeze(1).expand(-1, c, -1), grad_out * weight[..., q].unsqueeze(1))
    return out
