"""Point-set layers ("rows": one row per point, channels contiguous) on the B200 kernels: chains of 1x1 convolutions
(SharedMLP conv+BN+ReLU, Conv1d+bias+ReLU) as tcgen05 GEMMs with fused BN / activation passes and hand-written
backward; the set-abstraction scale (layer 0 on the points + gather, layers 1-2 on the grouped rows, max over neighbours)
and the three-NN interpolation.

Replaces, for the reference: pytorch_utils.SharedMLP on (B,C,npoint,nsample) tensors (cuDNN 1x1 convs + BN + ReLU +
F.max_pool2d, pointnet2_modules.py:60-69), grouping_operation / three_interpolate on channel-first tensors, and the
nn.Conv1d(k=1) stacks of ist_net.py:125-332.
"""
import ctypes
import os

import torch

from . import _C
from . import nhwc as K
from . import trace
from ._C import c_int, c_ll, ptr
from .nhwc import ACT_NONE, ACT_RELU, Act, ConvUnit


PN_NSPLIT = int(os.environ.get("ISTNET_NSPLIT_PN", "3"))  # forward operand planes of the PointNet++ SharedMLP GEMMs


def units_from_shared_mlp(mlp):
    """pytorch_utils.SharedMLP (conv1x1 no bias -> BN2d -> ReLU) * n  ->  [ConvUnit]"""
    return [ConvUnit(layer.conv.weight, None, layer.normlayer.bn, ACT_RELU, k=1, nsplit=PN_NSPLIT) for layer in mlp]


def units_from_conv1d_seq(seq, nsplit=None):
    """nn.Sequential of Conv1d(k=1) [+ ReLU] pairs -> [ConvUnit] (bias, optional ReLU)"""
    mods = [m for m in seq if not isinstance(m, torch.nn.AdaptiveAvgPool1d)]
    units = []
    for i, m in enumerate(mods):
        if isinstance(m, torch.nn.Conv1d):
            relu = i + 1 < len(mods) and isinstance(mods[i + 1], torch.nn.ReLU)
            units.append(ConvUnit(m.weight, m.bias, None, ACT_RELU if relu else ACT_NONE, k=1, nsplit=nsplit, nsplit_out=nsplit))
    return units


def _unit_params(units):
    ps = []
    for u in units:
        ps.append(u.w)
        if u.b is not None:
            ps.append(u.b)
        if u.bn is not None:
            ps += [u.bn.weight, u.bn.bias]
    return ps


def _chain_forward(units, a, training, record, last_f32=True, last_pair=False):
    tape = []
    for i, u in enumerate(units):
        last = i + 1 == len(units)
        a, rec = u.forward(a, training, record, want_f32=last and last_f32, want_pair=(not last) or last_pair)
        tape.append(rec)
    return a, tape


def _fusable_below(u, rec):
    return (K.FUSE_RELU_BWD and u.is_bias_relu(rec)) or (K.FUSE_BN_BWD and u.is_bn_relu(rec))


def _chain_backward(units, tape, dz, grads, need_dx_first, dy_first=None, below_first=None):
    """Backward through a chain of units.  dz: FP32 gradient w.r.t. the chain output (or dy_first: the top unit's dy operand
    planes, already made).  Wherever the layer below is bias+ReLU or BN+ReLU its activation backward rides in the epilogue of
    the data-gradient GEMM above it.  below_first = (unit, rec) of such a layer feeding units[0] from outside the chain: the
    return value is then that layer's FP32 dy instead of the gradient w.r.t. the chain input."""
    dy = dy_first  # operand planes of unit i's dy when the layer above already produced them in its data-gradient epilogue
    for i in range(len(units) - 1, -1, -1):
        u, rec = units[i], tape[i]
        if dy is None:
            dy, _ = u.act_backward(rec, dz, None, False, grads)
        plain = rec["kk"] == u.k and u.stride == 1
        if i > 0 and plain and _fusable_below(units[i - 1], tape[i - 1]) and tape[i - 1]["z_hi"].shape[-1] >= rec["xin"].C:
            dy, dz = u.data_grads(rec, dy, True, grads, below=(units[i - 1], tape[i - 1])), None
        elif i == 0 and below_first is not None and plain:
            return u.data_grads(rec, dy, True, grads, below=below_first, below_f32=True)
        else:
            dz, dy = u.data_grads(rec, dy, i > 0 or need_dx_first, grads), None
    return dz


def _split_sources(srcs, dev, nsplit=None):
    """[rows, c_i] FP32 tensors -> ONE operand-plane buffer [NSPLIT,1,1,rows,pad8(sum c_i)]: the channel concatenation of the
    reference (torch.cat at ist_net.py:168,172,255,323 ...) happens inside the split pass, no FP32 concat tensor is written."""
    rows = srcs[0].shape[0]
    cin = sum(t.shape[1] for t in srcs)
    a = Act(1, 1, rows, cin, srcs[0] if len(srcs) == 1 else None)
    a.pl = K.empty_planes(1, 1, rows, cin, dev, nsplit=nsplit)
    off = 0
    for t in srcs:
        K.split(t, rows, t.shape[1], a.pl, ch_off=off)
        off += t.shape[1]
    return a


class _ChainFn(torch.autograd.Function):
    """x_1 .. x_n [rows, c_i] FP32 (concatenated along the channels) -> [rows, Cout] FP32 through a list of ConvUnits."""

    @staticmethod
    def forward(ctx, units, training, nsrc, *tensors):
        srcs, params = tensors[:nsrc], tensors[nsrc:]
        out, tape = _chain_forward(units, _split_sources(srcs, srcs[0].device, units[0].ns), training, True)
        ctx.units, ctx.tape, ctx.params = units, tape, params
        ctx.widths = [t.shape[1] for t in srcs]
        ctx.need = [t.requires_grad for t in srcs]
        return out.f32.view(srcs[0].shape[0], -1)

    @staticmethod
    def backward(ctx, dz):
        grads = {}
        trace.mark(f"chain bwd> rows={dz.shape[0]} cout={dz.shape[1]}")
        dx = _chain_backward(ctx.units, ctx.tape, dz.contiguous().view(1, 1, dz.shape[0], dz.shape[1]), grads, any(ctx.need))
        K.join_side_streams()
        trace.mark("chain bwd<")
        ctx.tape = None
        dsrc, off = [], 0
        for w, need in zip(ctx.widths, ctx.need):
            dsrc.append(dx.view(dx.shape[2], dx.shape[3])[:, off : off + w] if (need and dx is not None) else None)
            off += w
        return (None, None, None) + tuple(dsrc) + tuple(grads.get(id(p)) if p.requires_grad else None for p in ctx.params)


def run_chain(units, x, training):
    """Differentiable chain of 1x1-conv units on a row matrix x [rows, Cin] (contiguous FP32), or on the channel
    concatenation of a list of such matrices."""
    srcs = [t.contiguous() for t in (x if isinstance(x, (list, tuple)) else (x,))]
    params = _unit_params(units)
    if torch.is_grad_enabled() and (any(t.requires_grad for t in srcs) or any(p.requires_grad for p in params)):
        return _ChainFn.apply(units, training, len(srcs), *srcs, *params)
    out, _ = _chain_forward(units, _split_sources(srcs, srcs[0].device, units[0].ns), training, False)
    return out.f32.view(srcs[0].shape[0], -1)


# ------------------------------------------------------------------------------------------ set abstraction scale
class _SAScaleFn(torch.autograd.Function):
    """One (radius, nsample) scale of PointnetSAModuleMSG (pointnet2_modules.py:60-69) in channels-last form:
    grouped rows -> 3 x [1x1 conv GEMM, BN, ReLU] -> max over the nsample axis."""

    @staticmethod
    def forward(ctx, units, training, xyz, new_xyz, idx, feats, *params):
        out, saved = sa_scale_forward(units, training, xyz, new_xyz, idx, feats, True)
        ctx.units, ctx.saved, ctx.params = units, saved, params
        ctx.has_feats = feats is not None and feats.requires_grad
        ctx.shape = (xyz.shape[0], xyz.shape[1], new_xyz.shape[1], idx.shape[2], 0 if feats is None else feats.shape[2])
        ctx.idx, ctx.xyz, ctx.new_xyz = idx, xyz, new_xyz
        return out

    @staticmethod
    def backward(ctx, dz):
        units, (tape, argmax, G, ns) = ctx.units, ctx.saved
        B, N, M, _, C = ctx.shape
        grads = {}
        dz = dz.contiguous()
        last = units[-1]
        rec = tape[-1]
        Cl = last.cout
        st = rec["bn"]
        dy = K.empty_planes(1, 1, G * ns, Cl, dz.device, nsplit=K.NSPLIT_BWD)
        # max-pool routing + ReLU mask + BN backward in one reduce / apply pair (no materialised [rows, C] selection tensor)
        _, sg_f, sgx_f = K.bn_act_bwd(dz, None, rec["y"], G * ns, Cl, 1, st, K.ACT_RELU_MAXROWS, None, None, None, dy_pl=dy, argmax=argmax, ns=ns)
        grads[id(last.bn.weight)], grads[id(last.bn.bias)] = sgx_f, sg_f
        d_feats = None
        if tape[0].get("point_l0"):
            # layers 1.. (the last one's dy is made above): every BN+ReLU layer's reduction rides in the data-gradient GEMM above it,
            # down to layer 0, whose FP32 dy comes out of layer 1's data-gradient GEMM + apply pass
            fuse0 = K.FUSE_BN_BWD and units[0].is_bn_relu(tape[0])
            d = _chain_backward(units[1:], tape[1:], None, grads, True, dy_first=dy, below_first=(units[0], tape[0]) if fuse0 else None)
            d_feats = _sa_l0_backward(units[0], tape[0], d, ctx.xyz, ctx.new_xyz, ctx.idx, ctx.has_feats, grads, have_dy0=fuse0)
        else:
            d = last.data_grads(rec, dy, True, grads)
            d = _chain_backward(units[:-1], tape[:-1], d, grads, ctx.has_feats)
            if ctx.has_feats:
                d_feats = torch.empty(B, N, C, dtype=torch.float32, device=dz.device)
                _C.call("group_rows_bwd", c_int(B), c_int(N), c_int(M), c_int(ns), c_int(C), ptr(d), ptr(ctx.idx), ptr(d_feats))
        K.join_side_streams()
        ctx.saved = None
        return (None, None, None, None, None, d_feats) + tuple(grads.get(id(p)) if p.requires_grad else None for p in ctx.params)


# Layer 0 of a set-abstraction scale on the points instead of the grouped rows (csrc/elementwise.cu, sa_gather_l0_kernel);
# ISTNET_SA_POINT_L0=0 keeps the round-1 dataflow (materialised grouped operand + grouped GEMM).
SA_POINT_L0 = os.environ.get("ISTNET_SA_POINT_L0", "1") != "0"


def _sa_l0_forward(u0, training, xyz, new_xyz, idx, feats, record):
    """y0 = Wx (xyz_j - c_i) + Wf f_j for every grouped row, BN (batch statistics from the gather kernel's partials) + ReLU,
    written as the operand planes of layer 1.  Returns (Act with planes, tape record)."""
    B, N, _ = xyz.shape
    M, ns = idx.shape[1], idx.shape[2]
    C = 0 if feats is None else feats.shape[2]
    rows, C0, dev = B * M * ns, u0.cout, xyz.device
    fa = uf = wf = None
    if C > 0:  # u = F Wf^T over the B*N points: 16..32x fewer rows than the grouped tensor
        fa = Act(1, 1, B * N, C, None, K.empty_planes(1, 1, B * N, C, dev))
        K.split(feats, B * N, C, fa.pl)
        wf = u0.w.detach()[:, 3:, 0, 0].contiguous()
        uf = torch.empty(1, 1, B * N, C0, dtype=torch.float32, device=dev)
        K.conv_gemm(fa, K.prep_weight(wf), C0, 1, 1, out_f32=uf)
    y0 = torch.empty(1, 1, rows, C0, dtype=torch.float32, device=dev)
    st, fin = K.bn_begin(u0.bn, C0, training, rows, dev)  # train mode: the gather kernel's last CTA finishes the statistics
    grid = ctypes.c_int(0)
    _C.call("sa_gather_l0", c_int(B), c_int(N), c_int(M), c_int(ns), c_int(C0), ptr(xyz), ptr(new_xyz), ptr(idx), K._p(uf), ptr(u0.w), c_int(3 + C),
            ptr(y0), ptr(K.stat_scratch(C0, dev)), ctypes.byref(grid), ctypes.byref(fin) if fin is not None else K.NULL)
    a = Act(1, 1, rows, C0, None, K.empty_planes(1, 1, rows, C0, dev))
    K.bn_act_split(y0, rows, C0, 1, bn=st, act=ACT_RELU, out_pl=a.pl)
    rec = {"point_l0": True, "bn": st}
    if record:
        rec.update({"y": y0, "z_hi": a.hi, "fa": fa, "wf": wf})
    return a, rec


def _sa_l0_backward(u0, rec, d, xyz, new_xyz, idx, need_dfeats, grads, have_dy0=False):
    """d: FP32 gradient w.r.t. layer 0's output rows.  BN/ReLU backward on the rows, then ONE pass scatters dy0 to the
    points (dU) and reduces dWx; the feature part of the weight gradient and the feature gradient are GEMMs over the
    points.  Returns d_feats (B,N,C) or None."""
    B, N, _ = xyz.shape
    M, ns = idx.shape[1], idx.shape[2]
    fa = rec["fa"]
    C = 0 if fa is None else fa.C
    rows, C0, dev = B * M * ns, u0.cout, d.device
    if have_dy0:  # d already is dy0 (BN/ReLU backward fused into layer 1's data-gradient GEMM, BN gradients set there)
        dy0 = d
    else:
        dy0 = torch.empty(1, 1, rows, C0, dtype=torch.float32, device=dev)
        _, sg_f, sgx_f = K.bn_act_bwd(d, None, rec["y"], rows, C0, 1, rec["bn"], ACT_RELU, None, rec["z_hi"], None, dy_f32=dy0)
        grads[id(u0.bn.weight)], grads[id(u0.bn.bias)] = sgx_f, sg_f
    dU = torch.empty(B * N, C0, dtype=torch.float32, device=dev) if C > 0 else None
    part = torch.empty(_C.lib().istnet_reduce_ws_floats(c_ll(rows), C0, 3), dtype=torch.float32, device=dev)
    wsx = torch.empty(3 * C0, dtype=torch.float64, device=dev)
    _C.call("sa_scatter_l0", c_int(B), c_int(N), c_int(M), c_int(ns), c_int(C0), ptr(dy0), ptr(xyz), ptr(new_xyz), ptr(idx), K._p(dU), ptr(part), ptr(wsx),
            _C.tickets(dev))
    gw = wsx.view(3, C0).t().float()
    d_feats = None
    if C > 0:
        dupl = K.empty_planes(1, 1, B * N, C0, dev, nsplit=K.NSPLIT_BWD)
        K.split(dU, B * N, C0, dupl)
        gw = torch.cat([gw, K.conv_wgrad(dupl, C0, fa, 1, 1).view(C0, C)], 1)
        if need_dfeats:
            d_feats = torch.empty(B, N, C, dtype=torch.float32, device=dev)
            K.conv_gemm(Act(1, 1, B * N, C0, None, dupl), K.prep_weight(rec["wf"], transpose=True, nsplit=dupl.shape[0]), C, 1, 1,
                        out_f32=d_feats.view(1, 1, B * N, C))
    grads[id(u0.w)] = gw.reshape(u0.w.shape)
    return d_feats


def sa_scale_forward(units, training, xyz, new_xyz, idx, feats, record):
    B, N, _ = xyz.shape
    M, ns = idx.shape[1], idx.shape[2]
    C = 0 if feats is None else feats.shape[2]
    rows = B * M * ns
    dev = xyz.device
    if SA_POINT_L0 and len(units) >= 2 and units[0].cout % 4 == 0:
        a, rec0 = _sa_l0_forward(units[0], training, xyz, new_xyz, idx, feats, record)
        a, tape = _chain_forward(units[1:-1], a, training, record, last_f32=False, last_pair=True)
        tape = [rec0] + tape
    else:
        a = Act(1, 1, rows, 3 + C)
        a.pl = K.empty_planes(1, 1, rows, 3 + C, dev)
        _C.call("group_rows_split", c_int(B), c_int(N), c_int(M), c_int(ns), c_int(C), ptr(xyz), ptr(new_xyz), K._p(feats), ptr(idx), *K._pl_args(a.pl),
                c_int(a.cs))
        a, tape = _chain_forward(units[:-1], a, training, record, last_f32=False, last_pair=True)
    last = units[-1]
    y, rec = last.forward(a, training, record, defer_act=True)
    tape.append(rec)
    st = rec["bn"]
    Cl = last.cout
    out = torch.empty(B, M, Cl, dtype=torch.float32, device=dev)
    argmax = torch.empty(B * M, Cl, dtype=torch.uint8, device=dev)
    _C.call("bn_relu_maxrows", ptr(y.f32), c_ll(B * M), c_int(ns), c_int(Cl), ptr(st.mean), ptr(st.invstd), ptr(st.gamma), ptr(st.beta), ptr(out),
            c_int(Cl), c_int(0), ptr(argmax))
    return out, (tape, argmax, B * M, ns)


def sa_scale(units, training, xyz, new_xyz, idx, feats):
    """xyz (B,N,3), new_xyz (B,M,3), idx int32 (B,M,ns), feats (B,N,C) channels-last or None -> (B,M,Cout)"""
    params = _unit_params(units)
    if feats is not None:
        feats = feats.contiguous()
    if torch.is_grad_enabled() and ((feats is not None and feats.requires_grad) or any(p.requires_grad for p in params)):
        return _SAScaleFn.apply(units, training, xyz, new_xyz, idx, feats, *params)
    out, _ = sa_scale_forward(units, training, xyz, new_xyz, idx, feats, False)
    return out


# ------------------------------------------------------------------------------------------ three-NN interpolation on rows
class _InterpRowsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feats, idx, weight):
        B, m, C = feats.shape
        n = idx.shape[1]
        out = torch.empty(B, n, C, dtype=torch.float32, device=feats.device)
        _C.call("interp_rows", c_int(B), c_int(m), c_int(n), c_int(C), ptr(feats), ptr(idx), ptr(weight), ptr(out), c_int(C), c_int(0))
        ctx.save_for_backward(idx, weight)
        ctx.m = m
        return out

    @staticmethod
    def backward(ctx, dout):
        idx, weight = ctx.saved_tensors
        dout = dout.contiguous()
        B, n, C = dout.shape
        d = torch.empty(B, ctx.m, C, dtype=torch.float32, device=dout.device)
        _C.call("interp_rows_bwd", c_int(B), c_int(ctx.m), c_int(n), c_int(C), ptr(dout), c_int(C), c_int(0), ptr(idx), ptr(weight), ptr(d))
        return d, None, None


def interp_rows(feats, idx, weight):
    """three_interpolate on channels-last features: feats (B,m,C), idx/weight (B,n,3) -> (B,n,C)"""
    return _InterpRowsFn.apply(feats.contiguous(), idx, weight)


# ------------------------------------------------------------------------------------------ feature-propagation level
class _FPLevelFn(torch.autograd.Function):
    """PointnetFPModule.forward (pointnet2_modules.py:164-209) on rows: three_nn + interpolation weights (one launch), three_interpolate
    + torch.cat([interpolated, skip]) + operand split (one launch), SharedMLP chain.  Round 1 issued three_nn, five ATen launches
    for the weights, interp_rows, two split passes and two .contiguous() copies for the same work."""

    @staticmethod
    def forward(ctx, units, training, unknown, known, known_rows, skip_rows, *params):
        B, n, _ = unknown.shape
        m, C2 = known.shape[1], known_rows.shape[2]
        C1 = 0 if skip_rows is None else skip_rows.shape[2]
        dev = unknown.device
        idx = torch.empty(B, n, 3, dtype=torch.int32, device=dev)
        w = torch.empty(B, n, 3, dtype=torch.float32, device=dev)
        _C.call("three_nn_weights", c_int(B), c_int(n), c_int(m), ptr(unknown), ptr(known), K.NULL, ptr(idx), ptr(w))
        a = Act(1, 1, B * n, C2 + C1)
        a.pl = K.empty_planes(1, 1, B * n, C2 + C1, dev, nsplit=units[0].ns)
        _C.call("interp_concat_split", c_int(B), c_int(m), c_int(n), c_int(C2), c_int(C1), ptr(known_rows), ptr(idx), ptr(w), K._p(skip_rows),
                *K._pl_args(a.pl), c_int(a.cs))
        out, tape = _chain_forward(units, a, training, not isinstance(ctx, _NoCtx))
        ctx.units, ctx.tape, ctx.params = units, tape, params
        ctx.idx, ctx.w, ctx.dims = idx, w, (B, m, n, C2, C1)
        ctx.need = (known_rows.requires_grad, skip_rows is not None and skip_rows.requires_grad)
        return out.f32.view(B, n, -1)

    @staticmethod
    def backward(ctx, dz):
        B, m, n, C2, C1 = ctx.dims
        grads = {}
        trace.mark(f"fp bwd> rows={B * n} cout={dz.shape[-1]}")
        dz = dz.contiguous()
        dx = _chain_backward(ctx.units, ctx.tape, dz.view(1, 1, B * n, dz.shape[-1]), grads, any(ctx.need))
        K.join_side_streams()
        d_known = d_skip = None
        if ctx.need[0]:
            d_known = torch.empty(B, m, C2, dtype=torch.float32, device=dz.device)
            _C.call("interp_rows_bwd", c_int(B), c_int(m), c_int(n), c_int(C2), ptr(dx), c_int(C2 + C1), c_int(0), ptr(ctx.idx), ptr(ctx.w), ptr(d_known))
        if ctx.need[1]:
            d_skip = dx.view(B, n, C2 + C1)[:, :, C2:]
        trace.mark("fp bwd<")
        ctx.tape = None
        return (None, None, None, None, d_known, d_skip) + tuple(grads.get(id(p)) if p.requires_grad else None for p in ctx.params)


def fp_level(units, training, unknown, known, known_rows, skip_rows):
    """unknown (B,n,3), known (B,m,3), known_rows (B,m,C2), skip_rows (B,n,C1) or None -> (B,n,Cout)"""
    params = _unit_params(units)
    known_rows = known_rows.contiguous()
    if skip_rows is not None:
        skip_rows = skip_rows.contiguous()
    C2, C1 = known_rows.shape[2], 0 if skip_rows is None else skip_rows.shape[2]
    if (C2 % 4) or (C1 % 4):
        return None  # caller falls back to the unfused flow
    tensors = (known_rows, skip_rows) if skip_rows is not None else (known_rows,)
    if torch.is_grad_enabled() and (any(t.requires_grad for t in tensors) or any(p.requires_grad for p in params)):
        return _FPLevelFn.apply(units, training, unknown.contiguous(), known.contiguous(), known_rows, skip_rows, *params)
    with torch.no_grad():
        return _FPLevelFn.forward(_NoCtx(), units, training, unknown.contiguous(), known.contiguous(), known_rows, skip_rows, *params)


class _NoCtx:
    """Stand-in for an autograd ctx when a Function's forward is used without recording."""
