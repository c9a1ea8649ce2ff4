"""The image branch (ResNet-18/8 + PSP + 3 up-sampling stages + head, model/modules.py:51-81, model/resnet.py:182-202)
executed channels-last on the B200 kernels with an explicit tape and a hand-written backward.

Forward dataflow (train): rgb -> im2col -> conv1 GEMM -> [BN+ReLU+MaxPool fused] -> 8 BasicBlocks (each conv is an
implicit-GEMM tcgen05 kernel; BN statistics in its epilogue, BN-apply + residual + ReLU + operand split one pass) -> PSP:
pyramid pooling in one pass, both 1x1 convolutions on the <= 36-pixel level maps, one pass for all up-samplings + their sum,
which enters the K = 512 bottleneck GEMM's activation pass as a residual -> [ReLU + Dropout2d scale] -> 3 x [bilinear x2
fused with the operand split -> conv3x3 GEMM -> BN + PReLU (+Dropout2d scale)] -> head: BN statistics of the final 1x1
convolution from the moments of its input, convolution + BN + PReLU only on the `choose`d rows (ist_net.py:41-45).
"""
import os

import torch
import torch.nn.functional as F

from . import _C
from . import nhwc as K
from . import trace
from ._C import c_int, c_ll, ptr
from .nhwc import ACT_NONE, ACT_PRELU, ACT_RELU, Act, ConvUnit


def _parse_ns_map(text):
    out = {}
    for item in text.split(","):
        if "=" in item:
            k, v = item.split("=")
            out[k.strip()] = int(v)
    return out


# Forward operand planes per layer group of the image branch ("prefix=planes,..." matched against the unit names of _units);
# 3 everywhere unless DESIGN.md section 2 documents a measured 2-plane choice
NS_MAP = _parse_ns_map(os.environ.get("ISTNET_NSPLIT_MAP", "up_=2,layer4=2"))


def _ns(name):
    for k, v in NS_MAP.items():
        if name.startswith(k):
            return v
    return None


def _units(net):
    """ConvUnit views of every conv of Modified_PSPNet, keyed like the state dict.  Rebuilt per call (cheap) so that
    module replicas (DataParallel) and re-materialised parameters are always the ones used."""
    f = net.feats
    u = {"conv1": ConvUnit(f.conv1.weight, None, f.bn1, ACT_RELU, k=7, stride=2, pad=3, nsplit=_ns("conv1"))}
    for li in (1, 2, 3, 4):
        for bi, blk in enumerate(getattr(f, f"layer{li}")):
            pre = f"layer{li}.{bi}"
            u[pre + ".conv1"] = ConvUnit(blk.conv1.weight, None, blk.bn1, ACT_RELU, k=3, stride=blk.stride, nsplit=_ns(pre))
            u[pre + ".conv2"] = ConvUnit(blk.conv2.weight, None, blk.bn2, ACT_RELU, k=3, nsplit=_ns(pre))
            if blk.downsample is not None:
                u[pre + ".down"] = ConvUnit(blk.downsample[0].weight, None, blk.downsample[1], ACT_NONE, k=1, stride=blk.stride, pad=0, nsplit=_ns(pre))
    for name in ("up_1", "up_2", "up_3"):
        seq = getattr(net, name).conv
        u[name] = ConvUnit(seq[1].weight, seq[1].bias, seq[2], ACT_PRELU, prelu=seq[3].weight, k=3, nsplit=_ns(name))
    u["final"] = ConvUnit(net.final[0].weight, net.final[0].bias, net.final[1], ACT_PRELU, prelu=net.final[2].weight, k=1)
    return u


def _basic_block_fwd(u, pre, x, training, record, tape):
    """resnet.py:50-66.  x: Act with f32 + pair."""
    c1, c2, dn = u[pre + ".conv1"], u[pre + ".conv2"], u.get(pre + ".down")
    z1, r1 = c1.forward(x, training, record, want_f32=False, want_pair=True)
    if dn is not None:
        yd, rd = dn.forward(x, training, record, defer_act=True)
        out, r2 = c2.forward(z1, training, record, res=yd.f32, res_bn=rd["bn"], want_f32=True)
    else:
        rd = None
        out, r2 = c2.forward(z1, training, record, res=x.f32, want_f32=True)
    if record:
        tape.append(("block", pre, r1, r2, rd))
    return out


def _basic_block_bwd(u, entry, dz, dz2, grads):
    """Backward of _basic_block_fwd.  dz (+dz2): gradient(s) w.r.t. the block output; returns the two gradient streams w.r.t. the
    block input (main path, identity / downsample path) — the block below adds them inside its own activation backward."""
    _, pre, r1, r2, rd = entry
    c1, c2, dn = u[pre + ".conv1"], u[pre + ".conv2"], u.get(pre + ".down")
    dmid, g = c2.backward(r2, dz, dz2, need_dx=True, g_out=True, grads=grads)
    dx1, _ = c1.backward(r1, dmid, None, need_dx=True, grads=grads)
    if dn is not None:
        dxd, _ = dn.backward(rd, g, None, need_dx=True, grads=grads)
        return dx1, dxd
    return dx1, g


def _stem_fwd(u, rgb, training, record, tape):
    """conv1 7x7/2 (im2col GEMM) + BN + ReLU + MaxPool(3,2,1) (resnet.py:182-186): rgb (B,3,H,W) NCHW -> Act [B,H/4,W/4,64]"""
    dev = rgb.device
    B, _, H, W = rgb.shape
    x0 = Act(B, H, W, 3)
    y0, r0 = u["conv1"].forward(x0, training, True, x_f32_nchw=rgb, defer_act=True)
    st0 = r0["bn"]
    Hp, Wp = (y0.H - 1) // 2 + 1, (y0.W - 1) // 2 + 1
    z = Act(B, Hp, Wp, 64, torch.empty(B, Hp, Wp, 64, dtype=torch.float32, device=dev))
    z.pl = K.empty_planes(B, Hp, Wp, 64, dev)
    argmax = torch.empty(B, Hp, Wp, 64, dtype=torch.uint8, device=dev)
    _C.call("bn_relu_maxpool", ptr(y0.f32), c_int(B), c_int(y0.H), c_int(y0.W), c_int(64), ptr(st0.mean), ptr(st0.invstd), ptr(st0.gamma),
            ptr(st0.beta), ptr(z.f32), *K._pl_args(z.pl), c_int(z.cs), ptr(argmax))
    if record:
        tape.append(("stem", r0, argmax, (y0.H, y0.W)))
    return z


def _stem_bwd(u, entry, dz, dz2, grads):
    _, r0, argmax, (H0, W0) = entry
    unit = u["conv1"]
    st = r0["bn"]
    B = r0["xin"].B
    dev = dz.device
    g0 = torch.empty(B, H0, W0, 64, dtype=torch.float32, device=dev)
    _C.call("maxpool_relu_bwd", ptr(r0["y"]), c_int(B), c_int(H0), c_int(W0), c_int(64), ptr(st.mean), ptr(st.invstd), ptr(st.gamma),
            ptr(st.beta), ptr(dz), K._p(dz2), ptr(argmax), ptr(g0))
    dy = K.empty_planes(B, H0, W0, 64, dev, nsplit=K.NSPLIT_BWD)
    _, sg_f, sgx_f = K.bn_act_bwd(g0, None, r0["y"], r0["P"], 64, H0 * W0, st, ACT_NONE, None, None, None, dy_pl=dy)
    grads[id(unit.bn.weight)], grads[id(unit.bn.bias)] = sgx_f, sg_f
    unit.data_grads(r0, dy, False, grads)


# ISTNET_DENSE_HEAD=1 keeps the round-1 head (dense 192x192x128 convolution output, statistics over it, dense BN backward)
DENSE_HEAD = os.environ.get("ISTNET_DENSE_HEAD", "0") == "1"


# ISTNET_PSP_KERNELS=0: pooled levels / up-sampling of the priors through ATen (adaptive_avg_pool2d, upsample_bilinear2d)
PSP_KERNELS = os.environ.get("ISTNET_PSP_KERNELS", "1") != "0"


def _sz4(sizes):
    return [c_int(v) for v in (list(sizes) + [0, 0, 0, 0])[:4]]


class _PspPool(torch.autograd.Function):
    """x [B,H,W,C] channels-last -> [B, sum s^2, C]: nn.AdaptiveAvgPool2d(s) for every pyramid level in one pass."""

    @staticmethod
    def forward(ctx, x, sizes):
        x = x.contiguous()
        B, H, W, C = x.shape
        out = torch.empty(B, sum(v * v for v in sizes), C, dtype=torch.float32, device=x.device)
        _C.call("psp_pool", ptr(x), c_int(B), c_int(H), c_int(W), c_int(C), *_sz4(sizes), ptr(out))
        ctx.sizes, ctx.shape = sizes, (B, H, W, C)
        return out

    @staticmethod
    def backward(ctx, d):
        B, H, W, C = ctx.shape
        dx = torch.empty(B, H, W, C, dtype=torch.float32, device=d.device)
        _C.call("psp_pool_bwd", ptr(d.contiguous()), c_int(B), c_int(H), c_int(W), c_int(C), *_sz4(ctx.sizes), ptr(dx))
        return dx, None


class _PspPrior(torch.autograd.Function):
    """t [B, sum s^2, C] -> [B,H,W,C]: sum over the levels of the bilinear (align_corners=False) up-sampling, one pass."""

    @staticmethod
    def forward(ctx, t, sizes, H, W):
        t = t.contiguous()
        B, _, C = t.shape
        out = torch.empty(B, H, W, C, dtype=torch.float32, device=t.device)
        _C.call("psp_prior", ptr(t), c_int(B), c_int(H), c_int(W), c_int(C), *_sz4(sizes), ptr(out))
        ctx.sizes, ctx.shape = sizes, (B, H, W, C, t.shape[1])
        return out

    @staticmethod
    def backward(ctx, g):
        B, H, W, C, cells = ctx.shape
        dt = torch.empty(B, cells, C, dtype=torch.float32, device=g.device)
        _C.call("psp_prior_bwd", ptr(g.contiguous()), c_int(B), c_int(H), c_int(W), c_int(C), *_sz4(ctx.sizes), ptr(dt))
        return dt, None, None, None


def _head_forward(unit, x, choose, training):
    """modules.py:64-66 + ist_net.py:42-45: y = W x + b on every pixel, train-mode BatchNorm over all B*H*W pixels, PReLU,
    then only the `choose`d pixels are used.  y is affine in x, so its batch statistics follow from the first two moments
    of x:  mean_y = W mean_x + b,  var_y[c] = w_c^T Cov(x) w_c  with  Cov = X^T X / P - mean_x mean_x^T.  X^T X is a
    64x64 tensor-core contraction over the pixels (the weight-gradient kernel applied to x against itself); the
    convolution itself runs on the B*N gathered rows only.  Returns (rows (B,N,128) FP32, record)."""
    dev = x.pl.device
    B, HW, C, Co = x.B, x.H * x.W, x.C, unit.cout
    P, N = B * HW, choose.shape[1]
    flat = (choose + torch.arange(B, device=dev, dtype=choose.dtype).view(B, 1) * HW).reshape(-1)
    xg = Act(1, 1, B * N, C, None, x.pl.view(x.pl.shape[0], P, x.cs).index_select(1, flat).view(x.pl.shape[0], 1, 1, B * N, x.cs))
    yg = torch.empty(1, 1, B * N, Co, dtype=torch.float32, device=dev)
    K.conv_gemm(xg, K.prep_weight(unit.w), Co, 1, 1, bias=unit.b, out_f32=yg)
    bn = unit.bn
    rec = {"x": x, "xg": xg, "yg": yg, "flat": flat, "P": P, "N": N}
    if K.bn_uses_batch_stats(bn, training):
        xpl = x.pl[: K.NSPLIT_BWD]  # two planes: the plane-rounding errors are unbiased and average out over ~1e6 pixels
        S = K.conv_wgrad(xpl, C, Act(x.B, x.H, x.W, C, None, xpl), 1, 1).view(C, C).double()
        part = torch.empty(_C.lib().istnet_reduce_ws_floats(c_ll(P), C, 1), dtype=torch.float32, device=dev)
        sx = torch.empty(C, dtype=torch.float64, device=dev)
        _C.call("colsum_planes", ptr(xpl), c_ll(xpl.stride(0)), c_int(xpl.shape[0]), c_ll(P), c_int(C), c_int(x.cs), ptr(part), ptr(sx))
        with torch.no_grad():
            W64, b64 = unit.w.view(Co, C).double(), unit.b.double()
            mx = sx / P
            cov = S / P - torch.outer(mx, mx)
            mean = W64 @ mx + b64
            var = ((W64 @ cov) * W64).sum(1).clamp_min_(0.0)
            invstd = torch.rsqrt(var + bn.eps)
            if bn.track_running_stats and bn.running_mean is not None:
                bn.num_batches_tracked += 1
                # momentum as a device scalar (nhwc.momentum_ptr): a scheduler update reaches a replayed CUDA graph;
                # momentum=None is nn.BatchNorm2d's cumulative average 1 / num_batches_tracked
                mom = K.momentum_tensor(bn, dev).double()
                mom = torch.where(mom < 0, 1.0 / bn.num_batches_tracked.double(), mom)
                bn.running_mean.copy_(((1.0 - mom) * bn.running_mean.double() + mom * mean).float())
                bn.running_var.copy_(((1.0 - mom) * bn.running_var.double() + mom * var * (P / max(P - 1, 1))).float())
        st = K.BnState(mean.float(), invstd.float(), bn.weight, bn.bias, batch=True)
        rec.update({"sx": sx, "S": S})
    else:
        st = K.BnState(bn.running_mean, torch.rsqrt(bn.running_var + bn.eps), bn.weight, bn.bias, batch=False)
    rec["bn"] = st
    out = torch.empty(B, N, Co, dtype=torch.float32, device=dev)
    K.bn_act_split(yg, B * N, Co, 1, bn=st, act=ACT_PRELU, prelu=unit.prelu, out_f32=out)
    return out, rec


def _head_backward(unit, rec, d_out, grads):
    """Backward of _head_forward.  With g = d_out * PReLU'(u) on the gathered rows (zero elsewhere) the train-mode BatchNorm
    backward  dy_p = gamma*invstd*(g_p - sum(g)/P - xhat_p*sum(g*xhat)/P)  is a sparse term plus a term AFFINE in x_p:
      dW = [sum_s (gamma*invstd*g_s) x_s^T] - a sx^T - diag(d) (W S + (b - mu) sx^T)
      dx_p = [W^T (gamma*invstd*g_p)] - A x_p - c,   A = W^T diag(d) W,  c = W^T (a + d*(b - mu))
    with a = gamma*invstd*sum(g)/P, d = gamma*invstd^2*sum(g*xhat)/P, S = X^T X, sx = sum_p x_p.  The bracketed terms are
    GEMMs over the B*N gathered rows; -A x - c is one 64->64 1x1 convolution over the feature map.  Returns dx (B,H,W,C)."""
    x, xg, yg, st, flat, P, N = rec["x"], rec["xg"], rec["yg"], rec["bn"], rec["flat"], rec["P"], rec["N"]
    dev = d_out.device
    C, Co, R = x.C, unit.cout, xg.W
    dys = K.empty_planes(1, 1, R, Co, dev, nsplit=K.NSPLIT_BWD)
    st_rows = K.BnState(st.mean, st.invstd, st.gamma, st.beta, batch=False)  # apply pass: dys = gamma*invstd*g (the sparse term)
    ws, sg_f, sgx_f = K.bn_act_bwd(d_out.view(1, 1, R, Co), None, yg, R, Co, 1, st_rows, ACT_PRELU, unit.prelu, None, None, dy_pl=dys)
    sg, sgx = ws[0:Co], ws[Co : 2 * Co]
    grads[id(unit.bn.weight)], grads[id(unit.bn.bias)] = sgx_f, sg_f
    grads[id(unit.prelu)] = ws[2 * Co : 3 * Co].sum().float().reshape(1)
    gw_s = K.conv_wgrad(dys, Co, Act(1, 1, R, C, None, xg.pl), 1, 1).view(Co, C)
    ds = torch.empty(1, 1, R, C, dtype=torch.float32, device=dev)
    K.conv_gemm(Act(1, 1, R, Co, None, dys), K.prep_weight(unit.w, transpose=True, nsplit=dys.shape[0]), C, 1, 1, out_f32=ds)
    if st.batch:
        with torch.no_grad():
            W64, b64 = unit.w.view(Co, C).double(), unit.b.double()
            gi = st.gamma.double() * st.invstd.double()
            a = gi * sg / P
            d = gi * st.invstd.double() * sgx / P
            bm = b64 - st.mean.double()
            sx, S = rec["sx"], rec["S"]
            gw = gw_s.double() - torch.outer(a, sx) - d[:, None] * (W64 @ S + torch.outer(bm, sx))
            A = W64.t() @ (d[:, None] * W64)
            c = W64.t() @ (a + d * bm)
        grads[id(unit.w)] = gw.float().view_as(unit.w)
        grads[id(unit.b)] = torch.zeros_like(unit.b)  # a bias feeding a train-mode BatchNorm has an identically zero gradient
        xb = x.pl[: K.NSPLIT_BWD]
        dx = torch.empty(x.B, x.H, x.W, C, dtype=torch.float32, device=dev)
        K.conv_gemm(Act(x.B, x.H, x.W, C, None, xb), K.prep_weight((-A).float().contiguous(), nsplit=xb.shape[0]), C, 1, 1,
                    bias=(-c).float().contiguous(), out_f32=dx)
    else:
        grads[id(unit.w)] = gw_s.view_as(unit.w)
        grads[id(unit.b)] = (st.gamma * st.invstd * sg.float()).detach()
        dx = torch.zeros(x.B, x.H, x.W, C, dtype=torch.float32, device=dev)
    dx.view(P, C).index_add_(0, flat, ds.view(R, C))  # gather backward: float atomics, as torch.gather's backward in the reference
    return dx


def forward(net, rgb, choose, training, record, u=None):
    """rgb (B,3,H,W) FP32 NCHW, choose (B,N) int64 -> rgb_local rows (B,N,128) FP32; tape (list) when record."""
    u = u or _units(net)
    dev = rgb.device
    B, _, H, W = rgb.shape
    tape = []
    rgb = rgb.contiguous()
    z = _stem_fwd(u, rgb, training, record, tape)
    trace.mark("image: stem done")
    # ---- layer1..4
    for li in (1, 2, 3, 4):
        for bi in (0, 1):
            z = _basic_block_fwd(u, f"layer{li}.{bi}", z, training, record, tape)
    trace.mark("image: layers done")
    # ---- PSP (modules.py:27-34).  Bilinear up-sampling acts per channel, so it commutes with the 1x1 bottleneck:
    #   bottleneck(cat(up(stage_i(feats)) ..., feats)) = Wb_x * feats + sum_i up(Wb_i * stage_i(feats)) + bias
    # The priors therefore stay on their tiny pooled maps (<= 36 pixels per instance) through BOTH 1x1 convolutions (torch
    # autograd sub-graph, a few hundred kFLOP), only their 1024-channel result is up-sampled, and the tensor-core GEMM
    # contracts K = 512 instead of 2560: the 2048-channel prior tensor (151 MB at B = 32), its concat / permute / operand
    # split, and 80 % of the bottleneck's forward, data-gradient and weight-gradient FLOPs are never executed.
    Hf, Wf = z.H, z.W
    wb = net.psp.bottleneck.weight  # [1024, 2560, 1, 1]: columns 512*i .. of stage i, the last 512 of feats (cat order)
    nst = len(net.psp.stages)
    cf = z.C
    sizes = [int(st[0].output_size[0] if isinstance(st[0].output_size, (tuple, list)) else st[0].output_size) for st in net.psp.stages]
    with torch.enable_grad() if record else torch.no_grad():
        leaf = z.f32.detach().requires_grad_(record)  # [B,Hf,Wf,512] channels-last
        if PSP_KERNELS and nst <= 4:
            # all pooled levels in one pass ([B, 1+4+9+36, 512]), the two 1x1 convolutions as tiny matmuls per level, all
            # up-samplings + their sum in one pass (csrc/elementwise.cu psp_*)
            pooled = _PspPool.apply(leaf, sizes)
            ts, off = [], 0
            for i, stage in enumerate(net.psp.stages):
                n = sizes[i] * sizes[i]
                rows = pooled[:, off : off + n].reshape(-1, cf)
                ts.append(((rows @ stage[1].weight.view(cf, cf).t()) @ wb[:, i * cf : (i + 1) * cf, 0, 0].t()).view(B, n, -1))
                off += n
            prior = _PspPrior.apply(torch.cat(ts, 1), sizes, Hf, Wf)
        else:
            nchw = leaf.permute(0, 3, 1, 2)
            prior = None
            for i, stage in enumerate(net.psp.stages):
                pooled = stage[0](nchw)  # AdaptiveAvgPool2d -> (B,512,s,s)
                sz = pooled.shape[-1]
                rows = pooled.permute(0, 2, 3, 1).reshape(-1, cf)
                t = (rows @ stage[1].weight.view(cf, cf).t()) @ wb[:, i * cf : (i + 1) * cf, 0, 0].t()  # (B*s*s, 1024)
                if sz == 1:
                    up = t.view(B, 1, 1, -1)  # a 1x1 map up-samples to a constant
                else:
                    up = F.interpolate(t.view(B, sz, sz, -1).permute(0, 3, 1, 2), size=(Hf, Wf), mode="bilinear", align_corners=False).permute(0, 2, 3, 1)
                prior = up if prior is None else prior + up
            prior = prior.expand(B, Hf, Wf, wb.shape[0]).contiguous()
    noise = _draw_noise(net, B, training, dev)
    ub = ConvUnit(wb[:, nst * cf :].detach().contiguous(), net.psp.bottleneck.bias, None, ACT_RELU, k=1, nsplit=_ns("bottleneck"))
    p, rb = ub.forward(z, training, record, noise=noise[0], res=prior.detach(), want_f32=True, want_pair=False)
    if record:
        tape.append(("psp", rb, leaf, prior, ub))
    trace.mark("image: psp done")
    # ---- up_1..3: bilinear x2 (align_corners=True) fused with the operand split, conv3x3, BN, PReLU, Dropout2d scale
    for i, name in enumerate(("up_1", "up_2", "up_3")):
        cin = p.C
        xu = Act(B, 2 * p.H, 2 * p.W, cin)
        xu.pl = K.empty_planes(B, xu.H, xu.W, cin, dev)
        K.upsample2x(p.f32, B, p.H, p.W, cin, xu.pl)
        last = name == "up_3"
        p, ru = u[name].forward(xu, training, record, noise=None if last else noise[i + 1], want_f32=not last, want_pair=last)
        if record:
            tape.append(("up", name, ru))
    trace.mark("image: ups done")
    if choose is None:  # dense feature map (ModifiedResnet.forward): the head on every pixel, no tape
        out, _ = u["final"].forward(p, training, False, want_f32=True, want_pair=False)
        return out.f32, None
    # ---- head: final 1x1 conv + BN + PReLU, evaluated only at the `choose`d pixels (see _head_forward)
    if DENSE_HEAD:
        yf, rf = u["final"].forward(p, training, True, defer_act=True)
        stf = rf["bn"]
        N = choose.shape[1]
        out = torch.empty(B, N, 128, dtype=torch.float32, device=dev)
        choose = choose.contiguous()
        _C.call("gather_bn_prelu", ptr(yf.f32), c_int(B), c_ll(yf.H * yf.W), c_int(128), c_int(N), ptr(choose), ptr(stf.mean), ptr(stf.invstd),
                ptr(stf.gamma), ptr(stf.beta), ptr(u["final"].prelu), ptr(out))
        if record:
            tape.append(("final", rf, choose, N))
    else:
        out, rh = _head_forward(u["final"], p, choose, training)
        if record:
            tape.append(("head", rh))
    return out, (tape if record else None)


def _draw_noise(net, B, training, dev):
    """Dropout2d masks (modules.py:56,62,72-78): bernoulli(1-p)/(1-p) per (b,c); tests may inject fixed masks."""
    if not training:
        return [None, None, None]
    out = []
    for c, p in ((1024, net.drop_1.p), (256, net.drop_2.p), (64, net.drop_2.p)):
        if p == 0.0:
            out.append(None)
        elif net.dropout_noise_fn is not None:
            out.append(net.dropout_noise_fn(B, c, p).to(dev).reshape(B, c).float().contiguous())
        else:
            out.append(torch.empty(B, c, device=dev, dtype=torch.float32).bernoulli_(1 - p).div_(1 - p))
    return out


def backward(net, tape, d_out, u):
    """d_out (B,N,128) -> {id(parameter): gradient} for every parameter of the branch that receives one."""
    grads = {}
    dev = d_out.device
    d_out = d_out.contiguous()
    dz, dz2 = None, None
    pending_wb = None
    for entry in reversed(tape):
        kind = entry[0]
        trace.mark("image bwd: " + kind + ("" if kind in ("final", "head", "psp", "stem") else " " + str(entry[1])))
        if kind == "final":
            _, rf, choose, N = entry
            unit = u["final"]
            xin, st = rf["xin"], rf["bn"]
            B, HW = xin.B, xin.H * xin.W
            g = torch.empty(B, xin.H, xin.W, 128, dtype=torch.float32, device=dev)
            slope = torch.empty(128, dtype=torch.float64, device=dev)
            _C.call("gather_bn_prelu_bwd", ptr(rf["y"]), c_int(B), c_ll(HW), c_int(128), c_int(N), ptr(choose), ptr(st.mean), ptr(st.invstd),
                    ptr(st.gamma), ptr(st.beta), ptr(unit.prelu), ptr(d_out), ptr(g), ptr(slope))
            dy = K.empty_planes(B, xin.H, xin.W, 128, dev, nsplit=K.NSPLIT_BWD)
            ws, sg_f, sgx_f = K.bn_act_bwd(g, None, rf["y"], rf["P"], 128, HW, st, ACT_NONE, None, None, None, dy_pl=dy)
            grads[id(unit.bn.weight)], grads[id(unit.bn.bias)] = sgx_f, sg_f
            grads[id(unit.b)] = torch.zeros_like(unit.b) if st.batch else (st.gamma * st.invstd * sg_f).detach()
            grads[id(unit.prelu)] = slope.sum().float().reshape(1)
            dz = unit.data_grads(rf, dy, True, grads)
            dz2 = None
        elif kind == "head":
            dz, dz2 = _head_backward(u["final"], entry[1], d_out, grads), None
        elif kind == "up":
            _, name, ru = entry
            dxu, _ = u[name].backward(ru, dz, dz2, need_dx=True, grads=grads)
            xin = ru["xin"]
            dz = K.upsample2x_bwd(dxu, xin.B, xin.H // 2, xin.W // 2, xin.C)
            dz2 = None
        elif kind == "psp":
            _, rb, leaf, prior, ub = entry
            # g = gradient w.r.t. (GEMM output + priors): the GEMM's dy and the priors' incoming gradient at once
            dxf, g = ub.backward(rb, dz, dz2, need_dx=True, g_out=True, grads=grads)
            stage_w = [s[1].weight for s in net.psp.stages]
            wb = net.psp.bottleneck.weight
            gs = torch.autograd.grad(prior, [leaf, wb] + stage_w, g)
            grads[id(wb)] = gs[1]
            pending_wb = (gs[1], len(stage_w) * leaf.shape[-1], id(ub.w))  # + the GEMM's weight gradient (side stream): added after the join
            for w_, g_ in zip(stage_w, gs[2:]):
                grads[id(w_)] = g_
            dz = dxf
            dz2 = gs[0].contiguous()  # leaf is channels-last: nothing to permute
        elif kind == "block":
            dz, dz2 = _basic_block_bwd(u, entry, dz, dz2, grads)
        elif kind == "stem":
            _stem_bwd(u, entry, dz, dz2, grads)
    K.join_side_streams()
    if pending_wb is not None:
        gwb, col0, key = pending_wb
        gwb[:, col0:] += grads.pop(key).reshape(gwb.shape[0], -1, 1, 1)
    return grads


class _ImageBranchFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, net, rgb, choose, *params):
        u = _units(net)
        out, tape = forward(net, rgb, choose, net.training, True, u)
        ctx.net, ctx.tape, ctx.params, ctx.units = net, tape, params, u
        return out

    @staticmethod
    def backward(ctx, d_out):
        trace.mark("image bwd>")
        grads = backward(ctx.net, ctx.tape, d_out, ctx.units)
        K.join_side_streams()
        trace.mark("image bwd<")
        ctx.tape = None
        return (None, None, None) + tuple(grads.get(id(p)) if p.requires_grad else None for p in ctx.params)


def image_branch(net, rgb, choose):
    """Modified_PSPNet forward + pixel gather on the B200 kernels -> rows (B,N,128); differentiable w.r.t. the parameters."""
    params = tuple(p for n, p in net.named_parameters() if not n.startswith("feats.fc."))
    if torch.is_grad_enabled() and any(p.requires_grad for p in params):
        return _ImageBranchFn.apply(net, rgb, choose, *params)
    out, _ = forward(net, rgb, choose, net.training, False)
    return out


def dense_map(net, rgb):
    """Modified_PSPNet.forward (modules.py:69-81) on the B200 kernels: (B,3,H,W) -> (B,128,H,W).  Forward only: the training path
    never needs the dense map (IST_Net reads it at the `choose`d pixels, ist_net.py:42-45)."""
    with torch.no_grad():
        out, _ = forward(net, rgb, None, net.training, False)
    return out.permute(0, 3, 1, 2).contiguous()
