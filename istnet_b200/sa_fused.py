"""Fused set-abstraction level on the kernels of csrc/sa_fused.cu: both (radius, nsample) scales of one
PointnetSAModuleMSG (pointnet2_modules.py:29-73,93-114) in ONE launch per pass — ball query + grouping + SharedMLP
(3 x [1x1 conv, train-mode BatchNorm, ReLU]) + max over nsample — with a hand-written backward.

Used for the two fine levels of PointNet2MSG (modules.py:249-275: MLPs [3|67] -> 16|32 -> 16|32 -> 32|64), whose grouped
tensors are large (0.8 M / 0.4 M rows per batch of 32) and whose layers are far too narrow for tensor-core tiles; the two
coarse levels keep the tcgen05 GEMM path of rows_engine.sa_scale.  Forward: [u GEMM on the points] + 3 passes + final =
4-5 launches per level (round 1: ~30); backward: pre + 3 stages [+ u backward] = 4-5 launches (round 1: ~60).
"""
import ctypes
import os

import torch

from . import _C
from . import nhwc as K
from ._C import c_int, c_void_p, ptr

ENABLED = os.environ.get("ISTNET_SA_FUSED", "1") != "0"


class SaScale(ctypes.Structure):
    """include/istnet_b200.h `istnet_sa_scale`"""

    _fields_ = [
        ("radius", ctypes.c_float), ("nsample", ctypes.c_int), ("idx", c_void_p), ("w0", c_void_p), ("ldw0", ctypes.c_int),
        ("w1", c_void_p), ("w2", c_void_p), ("bn_mean", c_void_p * 3), ("bn_invstd", c_void_p * 3), ("bn_gamma", c_void_p * 3),
        ("bn_beta", c_void_p * 3), ("part", c_void_p), ("fin", ctypes.POINTER(_C.Fin)), ("ysel", c_void_p), ("asel", c_void_p),
        ("dz", c_void_p), ("ld_dz", ctypes.c_int), ("off_dz", ctypes.c_int), ("ws2", c_void_p), ("ws1", c_void_p), ("ws0", c_void_p),
        ("g1", c_void_p), ("g0", c_void_p), ("part_w", c_void_p), ("tickets_w", c_void_p), ("dw", c_void_p), ("ld_dw", ctypes.c_int),
    ]


def _dp(t):
    return t.data_ptr() if t is not None else None


def supported(sa, xyz, new_xyz, feats):
    """Can this PointnetSAModuleMSG call run on the fused kernels?"""
    if not ENABLED or len(sa.mlps) != 2 or not xyz.is_cuda:
        return False
    if sorted(g.nsample for g in sa.groupers) != [16, 32] or sa.groupers[0].nsample != 16:
        return False
    C = 0 if feats is None else feats.shape[2]
    w = [[m[i].conv.weight.shape[0] for i in range(len(m))] for m in sa.mlps]
    if len(w[0]) != 3 or w[0] != w[1]:
        return False
    C0, C1, C2 = w[0]
    if not _C.lib().istnet_sa_level_supported(C0, C1, C2):
        return False
    if not ((C == 0) or (C == 64 and C0 == 32)):
        return False
    B, N, _ = xyz.shape
    M = new_xyz.shape[1]
    return B * M * 32 < 2**31 and B * N < 2**29 and N <= 8192


def _bn_of(mlp, l):
    return mlp[l].normlayer.bn


def _scales(sa, C, idx, states, dev):
    """SaScale array with the static part filled in (weights, ball parameters, BatchNorm pointers of `states[s][l]`)."""
    arr = (SaScale * 2)()
    for s in range(2):
        mlp, g = sa.mlps[s], sa.groupers[s]
        a = arr[s]
        a.radius, a.nsample, a.idx = float(g.radius), int(g.nsample), idx[s].data_ptr()
        a.w0, a.ldw0 = mlp[0].conv.weight.data_ptr(), 3 + C
        a.w1, a.w2 = mlp[1].conv.weight.data_ptr(), mlp[2].conv.weight.data_ptr()
        for l in range(3):
            st = states[s][l]
            if st is not None:
                a.bn_mean[l], a.bn_invstd[l] = st.mean.data_ptr(), st.invstd.data_ptr()
                a.bn_gamma[l], a.bn_beta[l] = st.gamma.data_ptr(), st.beta.data_ptr()
    return arr


def _forward(sa, training, xyz, new_xyz, feats):
    dev = xyz.device
    B, N, _ = xyz.shape
    M = new_xyz.shape[1]
    C = 0 if feats is None else feats.shape[2]
    C0, C1, C2 = (sa.mlps[0][i].conv.weight.shape[0] for i in range(3))
    widths = (C0, C1, C2)
    u = None
    if C > 0:  # feature part of layer 0 for both scales: one GEMM over the B*N points (16-32x fewer rows than the grouped tensor)
        u = torch.empty(B * N, 2 * C0, dtype=torch.float32, device=dev)
        _C.call("sa_u", c_int(B * N), c_int(C), c_int(C0), ptr(feats), ptr(sa.mlps[0][0].conv.weight), ptr(sa.mlps[1][0].conv.weight),
                c_int(3 + C), ptr(u))
    idx = [torch.empty(B, M, g.nsample, dtype=torch.int32, device=dev) for g in sa.groupers]
    ysel = [torch.empty(B * M, C2, dtype=torch.float32, device=dev) for _ in range(2)]
    asel = [torch.empty(B * M, C2, dtype=torch.uint8, device=dev) for _ in range(2)]
    states = [[None] * 3 for _ in range(2)]
    fins = [[None] * 3 for _ in range(2)]
    batch = K.bn_uses_batch_stats(_bn_of(sa.mlps[0], 0), training)
    for s in range(2):
        for l in range(3):
            bn = _bn_of(sa.mlps[s], l)
            assert K.bn_uses_batch_stats(bn, training) == batch, "mixed BatchNorm modes inside one set-abstraction level"
            states[s][l], fins[s][l] = K.bn_begin(bn, widths[l], training, B * M * sa.groupers[s].nsample, dev)
    dims = (c_int(B), c_int(N), c_int(M), c_int(C0), c_int(C1), c_int(C2), ptr(xyz), ptr(new_xyz), K._p(u), c_int(2 * C0))
    keep = []  # scratch referenced by the descriptors must outlive the launches (stream-ordered: the caching allocator handles reuse)
    for p in ((0, 1, 2) if batch else (2,)):
        arr = _scales(sa, C, idx, states, dev)
        for s in range(2):
            if batch:
                part = K.stat_scratch(widths[p], dev)
                keep.append(part)
                arr[s].part, arr[s].fin = part.data_ptr(), ctypes.pointer(fins[s][p])
            arr[s].ysel, arr[s].asel = ysel[s].data_ptr(), asel[s].data_ptr()
        _C.call("sa_level_forward", *dims, arr, c_int(p), c_int(1 if p == (0 if batch else 2) else 0))
    out = torch.empty(B, M, 2 * C2, dtype=torch.float32, device=dev)
    arr = _scales(sa, C, idx, states, dev)
    for s in range(2):
        arr[s].ysel = ysel[s].data_ptr()
    _C.call("sa_level_final", c_int(B), c_int(M), c_int(C2), arr, ptr(out), c_int(2 * C2))
    return out, (u, idx, ysel, asel, states)


class _SALevelFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sa, training, xyz, new_xyz, feats, *params):
        out, saved = _forward(sa, training, xyz, new_xyz, feats)
        ctx.sa, ctx.saved, ctx.params = sa, saved, params
        ctx.xyz, ctx.new_xyz, ctx.feats = xyz, new_xyz, feats
        ctx.need_dfeats = feats is not None and feats.requires_grad
        return out

    @staticmethod
    def backward(ctx, dz):
        sa, (u, idx, ysel, asel, states) = ctx.sa, ctx.saved
        xyz, new_xyz, feats = ctx.xyz, ctx.new_xyz, ctx.feats
        dev = dz.device
        dz = dz.contiguous()
        B, N, _ = xyz.shape
        M = new_xyz.shape[1]
        C = 0 if feats is None else feats.shape[2]
        C0, C1, C2 = (sa.mlps[0][i].conv.weight.shape[0] for i in range(3))
        widths = (C0, C1, C2)
        dims = (c_int(B), c_int(N), c_int(M), c_int(C0), c_int(C1), c_int(C2), ptr(xyz), ptr(new_xyz), K._p(u), c_int(2 * C0))
        f32 = dict(dtype=torch.float32, device=dev)
        ws = [[torch.empty(3 * widths[l], dtype=torch.float64, device=dev) for l in range(3)] for _ in range(2)]
        dgamma = [[torch.empty(widths[l], **f32) for l in range(3)] for _ in range(2)]
        dbeta = [[torch.empty(widths[l], **f32) for l in range(3)] for _ in range(2)]
        dW = [[torch.empty(C0, 3 + C, 1, 1, **f32), torch.empty(C1, C0, 1, 1, **f32), torch.empty(C2, C1, 1, 1, **f32)] for _ in range(2)]
        g1 = [torch.empty(B * M * sa.groupers[s].nsample, C1, **f32) for s in range(2)]
        g0 = [torch.empty(B * M * sa.groupers[s].nsample, C0, **f32) for s in range(2)]
        dU = torch.empty(B * N, 2 * C0, **f32) if C > 0 else None
        keep = []

        def launch(stage):
            # stage -1 finishes the sums of layer 2, stage 0 those of layer 1, stage 1 those of layer 0
            lsum = {-1: 2, 0: 1, 1: 0}.get(stage)
            lw = {0: 2, 1: 1, 2: 0}.get(stage)
            arr = _scales(sa, C, idx, states, dev)
            for s in range(2):
                a = arr[s]
                a.ysel, a.asel = ysel[s].data_ptr(), asel[s].data_ptr()
                a.dz, a.ld_dz, a.off_dz = dz.data_ptr(), 2 * C2, s * C2
                a.ws2, a.ws1, a.ws0 = ws[s][2].data_ptr(), ws[s][1].data_ptr(), ws[s][0].data_ptr()
                a.g1, a.g0 = g1[s].data_ptr(), g0[s].data_ptr()
                if lsum is not None:
                    part = K.stat_scratch(widths[lsum], dev, nacc=3)
                    fin = _C.Fin()
                    fin.kind, fin.tickets = _C.FIN_BN_BWD, _C.tickets(dev).value
                    fin.sum_f64, fin.sum_f32, fin.sum2_f32 = ws[s][lsum].data_ptr(), dbeta[s][lsum].data_ptr(), dgamma[s][lsum].data_ptr()
                    keep.extend((part, fin))
                    a.part, a.fin = part.data_ptr(), ctypes.pointer(fin)
                if lw is not None:
                    n = 3 * 32 if lw == 0 else dW[s][lw].numel()  # layer 0: [3 coordinates][32 lanes] partial rows
                    pw = torch.empty(_C.FIN_ROWS * n, **f32)
                    keep.append(pw)
                    a.part_w, a.tickets_w = pw.data_ptr(), _C.tickets(dev).value
                    a.dw, a.ld_dw = dW[s][lw].data_ptr(), dW[s][lw].shape[1]
            _C.call("sa_level_backward", *dims, K._p(dU if stage == 2 else None), arr, c_int(stage))

        for stage in (-1, 0, 1, 2):
            launch(stage)
        d_feats = None
        if C > 0:
            d_feats = torch.empty(B, N, C, **f32) if ctx.need_dfeats else None
            pw = torch.empty(_C.FIN_ROWS * 2 * C0 * C, **f32)
            dwf = torch.empty(2 * C0, C, **f32)
            _C.call("sa_u_bwd", c_int(B * N), c_int(C), c_int(C0), ptr(feats), ptr(dU), ptr(sa.mlps[0][0].conv.weight), ptr(sa.mlps[1][0].conv.weight),
                    c_int(3 + C), K._p(d_feats), ptr(pw), ptr(dwf))
            for s in range(2):  # layer-0 weight gradient = [dWx (3 columns, written by stage 2) | dWf]
                dW[s][0].view(C0, 3 + C)[:, 3:].copy_(dwf[s * C0 : (s + 1) * C0])
        grads = []
        for s in range(2):
            for l in range(3):
                grads += [dW[s][l], dgamma[s][l], dbeta[s][l]]
        ctx.saved = None
        return (None, None, None, None, d_feats) + tuple(g if p.requires_grad else None for g, p in zip(grads, ctx.params))


def level_params(sa):
    ps = []
    for s in range(2):
        for l in range(3):
            bn = _bn_of(sa.mlps[s], l)
            ps += [sa.mlps[s][l].conv.weight, bn.weight, bn.bias]
    return ps


def sa_level(sa, xyz, new_xyz, feats):
    """xyz (B,N,3), new_xyz (B,M,3), feats (B,N,C) channels-last or None -> (B,M,2*C2): both scales, radius-0 channels first."""
    params = level_params(sa)
    if feats is not None:
        feats = feats.contiguous()
    need_grad = torch.is_grad_enabled() and ((feats is not None and feats.requires_grad) or any(p.requires_grad for p in params))
    if need_grad:
        return _SALevelFn.apply(sa, sa.training, xyz, new_xyz, feats, *params)
    out, _ = _forward(sa, sa.training, xyz, new_xyz, feats)
    return out
