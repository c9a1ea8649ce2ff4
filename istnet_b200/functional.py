"""Autograd front-ends of the point-cloud kernels (autograd contract of SURVEY.md §8b: gradients flow to
feature tensors only, never to xyz / indices / weights — pointnet2_utils.py:72,110-114,145-146,180-203,235-254,283).
"""
import torch

from . import ext


class GroupPoints(torch.autograd.Function):
    """grouping_operation (pointnet2_utils.py:209-257): features (B,C,N), idx (B,m,ns) -> (B,C,m,ns)"""

    @staticmethod
    def forward(ctx, features, idx):
        ctx.n = features.shape[2]
        ctx.save_for_backward(idx)
        return ext.group_points(features.contiguous(), idx)

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        return ext.group_points_grad(grad_out.contiguous(), idx, ctx.n), None


class GatherPoints(torch.autograd.Function):
    """gather_operation (pointnet2_utils.py:83-117): features (B,C,N), idx (B,m) -> (B,C,m)"""

    @staticmethod
    def forward(ctx, features, idx):
        ctx.n = features.shape[2]
        ctx.save_for_backward(idx)
        return ext.gather_points(features.contiguous(), idx)

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        return ext.gather_points_grad(grad_out.contiguous(), idx, ctx.n), None


class ThreeInterpolate(torch.autograd.Function):
    """three_interpolate (pointnet2_utils.py:151-203): features (B,c,m), idx/weight (B,n,3) -> (B,c,n)"""

    @staticmethod
    def forward(ctx, features, idx, weight):
        ctx.m = features.shape[2]
        ctx.save_for_backward(idx, weight)
        return ext.three_interpolate(features.contiguous(), idx, weight)

    @staticmethod
    def backward(ctx, grad_out):
        idx, weight = ctx.saved_tensors
        return ext.three_interpolate_grad(grad_out.contiguous(), idx, weight, ctx.m), None, None


group_points = GroupPoints.apply
gather_points = GatherPoints.apply
three_interpolate = ThreeInterpolate.apply


@torch.no_grad()
def furthest_point_sample(xyz, npoint):
    return ext.furthest_point_sampling(xyz.contiguous(), npoint)


@torch.no_grad()
def ball_query(radius, nsample, xyz, new_xyz):
    """Argument order of the Python-level reference wrapper (pointnet2_utils.py:260-291)."""
    return ext.ball_query(new_xyz.contiguous(), xyz.contiguous(), radius, nsample)


@torch.no_grad()
def three_nn(unknown, known):
    """Returns (dist, idx) with dist = sqrt(dist2) as pointnet2_utils.py:142 does."""
    dist2, idx = ext.three_nn(unknown.contiguous(), known.contiguous())
    return torch.sqrt(dist2), idx


@torch.no_grad()
def three_nn_weights(unknown, known):
    """three_nn + inverse-distance weights of PointnetFPModule.forward (pointnet2_modules.py:185-188)."""
    dist, idx = three_nn(unknown, known)
    recip = 1.0 / (dist + 1e-8)
    return idx, (recip / torch.sum(recip, dim=2, keepdim=True)).contiguous()
