"""Tight, well-conditioned GPU checks of the individual dense / fused kernels (through the C ABI) against FLOAT64 PyTorch
evaluations of the reference ops they replace.  Tolerances are written per test; the float target of the path is 1e-4
relative (max|a-b| / max|b|)."""
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def K():
    from istnet_b200 import nhwc

    return nhwc


@pytest.fixture(scope="module")
def tc():
    from istnet_b200 import tc as t

    return t


CONV_CASES = [(2, 8, 8, 64, 64, 1), (2, 8, 8, 64, 64, 3), (2, 24, 24, 128, 256, 3), (3, 24, 24, 512, 512, 3), (1, 1, 1000, 320, 384, 1),
              (1, 1, 4096, 67, 32, 1), (1, 1, 300, 3, 16, 1), (2, 48, 48, 1024, 256, 3), (2, 16, 16, 128, 18, 1), (1, 1, 512, 2560, 1024, 1),
              (1, 8, 8, 512, 512, 3), (1, 1, 640, 256, 384, 1), (2, 8, 8, 96, 320, 3), (1, 1, 256, 64, 600, 1)]


@pytest.mark.parametrize("B,H,W,cin,cout,k", CONV_CASES)
def test_conv_gemm_matches_float64_conv2d(tc, B, H, W, cin, cout, k):
    """tcgen05 implicit GEMM (nn.Conv2d / Conv1d / Linear call sites) — 1e-4; odd batch, K tails, N tails, 1x1 and 3x3."""
    g = torch.Generator(device="cuda").manual_seed(B * 131 + cin + cout)
    x = torch.randn(B, cin, H, W, device="cuda", generator=g)
    w = torch.randn(cout, cin, k, k, device="cuda", generator=g) / (cin * k * k) ** 0.5
    b = torch.randn(cout, device="cuda", generator=g)
    ref = F.conv2d(x.double(), w.double(), b.double(), padding=k // 2).relu().permute(0, 2, 3, 1)
    ap = tc.split_planes_torch(x.permute(0, 2, 3, 1).contiguous())
    wp = tc.split_planes_torch(w.permute(2, 3, 0, 1).reshape(k * k, cout, cin).contiguous())
    out, planes = tc.conv_gemm(ap, cin, wp, cout, k, k, bias=b, relu=True, out_split=True)
    assert rel_err(out, ref) < 1e-4
    assert rel_err(planes.float().sum(0)[..., :cout], ref) < 1e-4  # fused operand-plane epilogue


@pytest.mark.parametrize("B,H,W,cin,cout,k", [(2, 8, 8, 64, 64, 1), (2, 8, 8, 64, 128, 3), (4, 24, 24, 128, 256, 3), (2, 24, 24, 512, 512, 3),
                                               (1, 1, 4096, 320, 384, 1), (1, 1, 1000, 67, 32, 1), (3, 16, 16, 128, 18, 1), (1, 1, 512, 3, 16, 1),
                                               # Cout <= 64 with several taps: two taps per accumulator tile (shifted dY), odd sizes, 25 taps
                                               (2, 12, 20, 64, 64, 3), (3, 9, 7, 256, 64, 3), (2, 16, 16, 32, 16, 3), (1, 10, 10, 64, 48, 3),
                                               (1, 12, 12, 16, 32, 5), (2, 48, 48, 64, 64, 3)])
def test_wgrad_matches_float64_autograd(tc, B, H, W, cin, cout, k):
    """tcgen05 weight gradient (MN-major operands, split-K, deterministic reduce) — 2e-5."""
    g = torch.Generator(device="cuda").manual_seed(cin * 7 + cout)
    x = torch.randn(B, cin, H, W, device="cuda", dtype=torch.float64, generator=g)
    w = torch.randn(cout, cin, k, k, device="cuda", dtype=torch.float64, generator=g).requires_grad_(True)
    y = F.conv2d(x, w, None, padding=k // 2)
    dy = torch.randn(y.shape, device="cuda", dtype=torch.float64, generator=g)
    (gw_ref,) = torch.autograd.grad(y, w, dy)
    xp = tc.split_planes_torch(x.float().permute(0, 2, 3, 1).contiguous())
    dp = tc.split_planes_torch(dy.float().permute(0, 2, 3, 1).contiguous())
    gw = tc.conv_wgrad(dp, cout, xp, cin, k, k)
    gw2 = tc.conv_wgrad(dp, cout, xp, cin, k, k)
    assert torch.equal(gw, gw2)  # deterministic
    assert rel_err(gw, gw_ref) < 2e-5


def _unit_vs_torch(K, B, H, W, cin, cout, k, stride, bn, act, bias, noise, seed):
    from istnet_b200.nhwc import Act, ConvUnit

    torch.manual_seed(seed)
    g = torch.Generator(device="cuda").manual_seed(seed)
    conv = torch.nn.Conv2d(cin, cout, k, stride=stride, padding=k // 2, bias=bias).cuda()
    bnm = torch.nn.BatchNorm2d(cout).cuda() if bn else None
    prelu = torch.nn.PReLU().cuda() if act == 2 else None
    # keep every pre-activation far from the PReLU kink (alternating +-8 offsets per channel): the derivative is
    # discontinuous there and a rounding-level sign flip would change single gradient elements by O(1)
    sign = (torch.arange(cout, device="cuda") % 2 * 2 - 1).float()
    if bnm is not None:
        torch.nn.init.uniform_(bnm.weight, 0.5, 1.0)
        with torch.no_grad():
            # ReLU (act 1): +-3 keeps a realistic share of masked elements per channel while making a rounding-level sign flip of
            # a pre-activation (which would change single gradient elements by O(1)) vanishingly unlikely
            bnm.bias.copy_(8.0 * sign if act == 2 else (3.0 * sign if act == 1 else 0.3 * torch.randn(cout, device="cuda")))
        bnm.momentum = 0.7
    elif bias and act != 0:
        with torch.no_grad():
            conv.bias.copy_(8.0 * sign)
    x = torch.randn(B, cin, H, W, device="cuda", generator=g)
    nz = (torch.rand(B, cout, device="cuda", generator=g) > 0.3).float() / 0.7 if noise else None
    # float64 reference
    import copy

    c64, b64, p64 = copy.deepcopy(conv).double(), copy.deepcopy(bnm).double() if bnm else None, copy.deepcopy(prelu).double() if prelu else None
    x64 = x.double().requires_grad_(True)
    u = c64(x64)
    if b64 is not None:
        u = b64.train()(u)
    z = u if act == 0 else (p64(u) if act == 2 else u.relu())
    if nz is not None:
        z = z * nz.double()[:, :, None, None]
    dz = torch.randn(z.shape, device="cuda", dtype=torch.float64, generator=g)
    z.backward(dz)
    # B200 unit
    unit = ConvUnit(conv.weight, conv.bias, bnm, act, prelu=prelu.weight if prelu else None, k=k, stride=stride)
    xin = Act(B, H, W, cin, x.permute(0, 2, 3, 1).contiguous())
    xin.pl = K.empty_planes(B, H, W, cin, "cuda")
    K.split(xin.f32, B * H * W, cin, xin.pl)
    out, rec = unit.forward(xin, True, True, noise=nz, want_f32=True)
    grads = {}
    dx, _ = unit.backward(rec, dz.float().permute(0, 2, 3, 1).contiguous(), None, need_dx=True, grads=grads)
    res = {"z": (out.f32, z.detach().permute(0, 2, 3, 1)), "dx": (dx, x64.grad.permute(0, 2, 3, 1)), "dw": (grads[id(conv.weight)], c64.weight.grad)}
    if bnm is not None:
        res["dgamma"] = (grads[id(bnm.weight)], b64.weight.grad)
        res["dbeta"] = (grads[id(bnm.bias)], b64.bias.grad)
        res["running_mean"] = (bnm.running_mean, b64.running_mean)
        res["running_var"] = (bnm.running_var, b64.running_var)
    elif bias:
        res["dbias"] = (grads[id(conv.bias)], c64.bias.grad)
    if prelu is not None:
        res["dslope"] = (grads[id(prelu.weight)], p64.weight.grad)
    return res


@pytest.mark.parametrize("cfg", [
    dict(B=4, H=16, W=16, cin=64, cout=128, k=3, stride=1, bn=True, act=0, bias=False, noise=False),   # conv + BN(train)
    dict(B=4, H=16, W=16, cin=64, cout=64, k=3, stride=1, bn=True, act=2, bias=True, noise=True),      # PSPUpsample tail
    dict(B=2, H=1, W=2048, cin=320, cout=256, k=1, stride=1, bn=False, act=2, bias=True, noise=False),  # per-point layer
    dict(B=4, H=16, W=16, cin=64, cout=128, k=3, stride=2, bn=True, act=0, bias=False, noise=False),   # strided (im2col) conv
    dict(B=4, H=16, W=16, cin=64, cout=128, k=1, stride=2, bn=True, act=0, bias=False, noise=False),   # strided downsample
    dict(B=4, H=16, W=16, cin=64, cout=128, k=3, stride=1, bn=True, act=1, bias=False, noise=False),   # conv + BN(train) + ReLU (every ResNet / SharedMLP layer)
    dict(B=2, H=1, W=4096, cin=128, cout=128, k=1, stride=1, bn=True, act=1, bias=False, noise=False),  # SharedMLP layer on rows
])
def test_conv_bn_act_unit_forward_backward(K, cfg):
    """conv (+bias) -> BatchNorm(train) -> PReLU -> Dropout2d scale, forward and hand-written backward, vs float64 autograd.
    Pre-activations are kept away from the activation kink (see _unit_vs_torch): 1e-4 on everything, including the BN
    running statistics."""
    res = _unit_vs_torch(K, seed=11, **cfg)
    for name, (a, b) in res.items():
        assert rel_err(a, b) < 1e-4, (name, rel_err(a, b))


def test_upsample2x_and_adjoint(K):
    """nn.Upsample(x2, bilinear, align_corners=True) fused with the operand split, and its gather-form adjoint — 1e-6."""
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn(3, 64, 12, 12, device="cuda", generator=g, dtype=torch.float64).requires_grad_(True)
    ref = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)
    d = torch.randn(ref.shape, device="cuda", generator=g, dtype=torch.float64)
    ref.backward(d)
    pl = K.empty_planes(3, 24, 24, 64, "cuda")
    K.upsample2x(x.detach().float().permute(0, 2, 3, 1).contiguous(), 3, 12, 12, 64, pl)
    assert rel_err(pl.float().sum(0), ref.detach().permute(0, 2, 3, 1)) < 1e-6
    dx = K.upsample2x_bwd(d.float().permute(0, 2, 3, 1).contiguous(), 3, 12, 12, 64)
    assert rel_err(dx, x.grad.permute(0, 2, 3, 1)) < 1e-6


def test_sa_scale_and_interp_rows_match_reference_flow():
    """Fused set-abstraction scale (group -> 3 x [GEMM, BN, ReLU] -> max) and row interpolation vs the reference dataflow
    (grouping_operation + SharedMLP + max_pool2d, three_interpolate) evaluated in float64 — forward 1e-4."""
    from istnet_b200 import ext, rows_engine as RE
    from istnet_b200.pointnet2 import SharedMLP
    from istnet_b200.synth import make_batch

    d = make_batch(4, 512, 8, seed=3)
    xyz = (d["pts"] - d["pts"].mean(1, keepdim=True)).cuda().contiguous()
    g = torch.Generator(device="cuda").manual_seed(1)
    feats = torch.randn(4, 512, 64, device="cuda", generator=g).requires_grad_(True)
    idx_f = ext.furthest_point_sampling(xyz, 128)
    new_xyz = torch.gather(xyz, 1, idx_f.long()[..., None].expand(-1, -1, 3)).contiguous()
    idx = ext.ball_query(new_xyz, xyz, 0.04, 32)
    mlp = SharedMLP([67, 32, 32, 64]).cuda().train()
    out = RE.sa_scale(RE.units_from_shared_mlp(mlp), True, xyz, new_xyz, idx, feats)
    m64 = SharedMLP([67, 32, 32, 64]).cuda().double().train()
    m64.load_state_dict(mlp.state_dict())
    gi = idx.long()
    gx = torch.gather(xyz.double()[:, None].expand(-1, 128, -1, -1), 2, gi[..., None].expand(-1, -1, -1, 3)) - new_xyz.double()[:, :, None]
    f64 = feats.detach().double().requires_grad_(True)
    gf = torch.gather(f64[:, None].expand(-1, 128, -1, -1), 2, gi[..., None].expand(-1, -1, -1, 64))
    grouped = torch.cat([gx, gf], -1).permute(0, 3, 1, 2)  # (B, 3+C, npoint, nsample)
    ref = F.max_pool2d(m64(grouped), kernel_size=[1, 32]).squeeze(-1).transpose(1, 2)
    assert rel_err(out, ref) < 1e-4
    # backward: layer 0 is evaluated on the points (u = F Wf^T gathered per row + Wx (xyz_j - c_i)); its gradients (scatter to
    # the points, point-level weight / feature GEMMs) must match autograd through the grouped reference flow
    cot = torch.randn(out.shape, device="cuda", generator=g)
    out.backward(cot)
    ref.backward(cot.double())
    assert rel_err(feats.grad, f64.grad) < 1e-4
    for (n, p), p64 in zip(mlp.named_parameters(), m64.parameters()):
        assert rel_err(p.grad, p64.grad) < 2e-4, n
    d2, i3 = ext.three_nn(xyz, new_xyz)
    w = torch.rand(4, 512, 3, device="cuda", generator=g)
    got = RE.interp_rows(out.detach().contiguous(), i3, w)
    want = sum(torch.gather(out.detach(), 1, i3[..., q].long()[..., None].expand(-1, -1, 64)) * w[..., q : q + 1] for q in range(3))
    assert rel_err(got, want) < 1e-6


def test_head_from_input_moments_matches_dense_float64(K):
    """image_engine._head_forward/_head_backward: 1x1 conv + train-mode BN + PReLU read only at the `choose`d pixels, with the BN
    statistics taken from sum(x) and X^T X and the BN backward split into sparse rows + one affine 64->64 convolution, against the
    dense float64 flow of the reference (modules.py:64-66, ist_net.py:42-45): outputs, running statistics, every gradient — 1e-4."""
    import copy

    from istnet_b200 import image_engine as IE
    from istnet_b200.nhwc import ACT_PRELU, Act, ConvUnit

    torch.manual_seed(5)
    g = torch.Generator(device="cuda").manual_seed(5)
    B, H, W, C, Co, N = 3, 24, 24, 64, 128, 200
    conv = torch.nn.Conv2d(C, Co, 1).cuda()
    bn = torch.nn.BatchNorm2d(Co).cuda().train()
    prelu = torch.nn.PReLU().cuda()
    sign = (torch.arange(Co, device="cuda") % 2 * 2 - 1).float()
    torch.nn.init.uniform_(bn.weight, 0.5, 1.0)
    with torch.no_grad():
        bn.bias.copy_(8.0 * sign)  # keep u away from the PReLU kink (see _unit_vs_torch)
    bn.momentum = 0.3
    x = (torch.randn(B, H, W, C, device="cuda", generator=g) * 0.7 + 0.4).relu()  # non-zero mean, like a PReLU output
    choose = torch.randint(0, H * W, (B, N), device="cuda", generator=g)
    choose[:, 1] = choose[:, 0]  # a pixel chosen twice: its gradient accumulates
    cot = torch.randn(B, N, Co, device="cuda", generator=g)
    # dense float64 reference
    c64, b64, p64 = copy.deepcopy(conv).double(), copy.deepcopy(bn).double(), copy.deepcopy(prelu).double()
    x64 = x.double().permute(0, 3, 1, 2).contiguous().requires_grad_(True)
    dense = p64(b64(c64(x64))).view(B, Co, H * W)
    ref = torch.gather(dense, 2, choose[:, None, :].expand(-1, Co, -1)).transpose(1, 2)
    ref.backward(cot.double())
    # B200 path
    unit = ConvUnit(conv.weight, conv.bias, bn, ACT_PRELU, prelu=prelu.weight, k=1)
    xa = Act(B, H, W, C, None, K.empty_planes(B, H, W, C, "cuda"))
    K.split(x.contiguous(), B * H * W, C, xa.pl)
    out, rec = IE._head_forward(unit, xa, choose, True)
    grads = {}
    dx = IE._head_backward(unit, rec, cot.contiguous(), grads)
    assert rel_err(out, ref) < 1e-4
    assert rel_err(bn.running_mean, b64.running_mean) < 1e-4 and rel_err(bn.running_var, b64.running_var) < 1e-4
    assert int(bn.num_batches_tracked) == 1
    assert rel_err(dx, x64.grad.permute(0, 2, 3, 1)) < 1e-4
    assert rel_err(grads[id(conv.weight)], c64.weight.grad) < 1e-4
    assert rel_err(grads[id(bn.weight)], b64.weight.grad) < 1e-4 and rel_err(grads[id(bn.bias)], b64.bias.grad) < 1e-4
    assert rel_err(grads[id(prelu.weight)], p64.weight.grad) < 1e-4
    assert float(grads[id(conv.bias)].abs().max()) == 0.0 and float(c64.bias.grad.abs().max()) < 1e-9 * float(cot.abs().sum())


def test_bias_relu_chain_backward_fused_into_dgrad_epilogue():
    """nn.Conv1d(k=1)+ReLU stack (ist_net.py:130-160) on rows_engine.run_chain: the interior layers' ReLU backward, bias gradient and
    dy operand split ride in the data-gradient GEMM epilogue above them (conv_gemm mask_hi).  Against float64 autograd: 1e-4."""
    import copy

    from istnet_b200 import nhwc, rows_engine as RE

    assert nhwc.FUSE_RELU_BWD
    torch.manual_seed(9)
    g = torch.Generator(device="cuda").manual_seed(9)
    seq = torch.nn.Sequential(torch.nn.Conv1d(64, 384, 1), torch.nn.ReLU(), torch.nn.Conv1d(384, 256, 1), torch.nn.ReLU(),
                              torch.nn.Conv1d(256, 18, 1)).cuda()
    s64 = copy.deepcopy(seq).double()
    x = torch.randn(1000, 64, device="cuda", generator=g).requires_grad_(True)
    x64 = x.detach().double().requires_grad_(True)
    out = RE.run_chain(RE.units_from_conv1d_seq(seq), x, True)
    ref = s64(x64.t()[None]).squeeze(0).t()
    assert rel_err(out, ref) < 1e-4
    cot = torch.randn(out.shape, device="cuda", generator=g)
    out.backward(cot)
    ref.backward(cot.double())
    assert rel_err(x.grad, x64.grad) < 1e-4
    for (n, p), q in zip(seq.named_parameters(), s64.parameters()):
        assert rel_err(p.grad, q.grad.float()) < 1e-4, n


@pytest.mark.parametrize("B,H,W,C,Co", [(3, 24, 24, 64, 128), (2, 8, 8, 32, 64), (2, 7, 10, 16, 16)])
def test_psp_pool_and_prior_kernels_match_aten(B, H, W, C, Co):
    """psp_pool == nn.AdaptiveAvgPool2d(s) for s in (1,2,3,6) (incl. the overlapping bins of maps not divisible by s), psp_prior ==
    sum of F.interpolate(bilinear, align_corners=False) of the level maps (modules.py:17,30), and both adjoints — vs float64 ATen."""
    from istnet_b200.image_engine import _PspPool, _PspPrior

    g = torch.Generator(device="cuda").manual_seed(3)
    sizes = [1, 2, 3, 6]
    x = torch.randn(B, H, W, C, device="cuda", generator=g).requires_grad_(True)
    x64 = x.detach().double().requires_grad_(True)
    pooled = _PspPool.apply(x, sizes)
    ref = torch.cat([F.adaptive_avg_pool2d(x64.permute(0, 3, 1, 2), s).permute(0, 2, 3, 1).reshape(B, s * s, C) for s in sizes], 1)
    assert rel_err(pooled, ref) < 1e-6
    cot = torch.randn(pooled.shape, device="cuda", generator=g)
    pooled.backward(cot)
    ref.backward(cot.double())
    assert rel_err(x.grad, x64.grad) < 1e-6
    t = torch.randn(B, sum(s * s for s in sizes), Co, device="cuda", generator=g).requires_grad_(True)
    t64 = t.detach().double().requires_grad_(True)
    prior = _PspPrior.apply(t, sizes, H, W)
    off, want = 0, 0
    for s in sizes:
        m = t64[:, off : off + s * s].view(B, s, s, Co).permute(0, 3, 1, 2)
        want = want + F.interpolate(m, size=(H, W), mode="bilinear", align_corners=False).permute(0, 2, 3, 1)
        off += s * s
    assert rel_err(prior, want) < 1e-6
    cot = torch.randn(prior.shape, device="cuda", generator=g)
    prior.backward(cot)
    want.backward(cot.double())
    assert rel_err(t.grad, t64.grad) < 1e-6


def _block_units(blk, pre):
    from istnet_b200.nhwc import ACT_NONE, ACT_RELU, ConvUnit

    u = {pre + ".conv1": ConvUnit(blk.conv1.weight, None, blk.bn1, ACT_RELU, k=3, stride=blk.stride),
         pre + ".conv2": ConvUnit(blk.conv2.weight, None, blk.bn2, ACT_RELU, k=3)}
    if blk.downsample is not None:
        u[pre + ".down"] = ConvUnit(blk.downsample[0].weight, None, blk.downsample[1], ACT_NONE, k=1, stride=blk.stride, pad=0)
    return u


@pytest.mark.parametrize("cin,cout,stride", [(64, 64, 1), (64, 128, 2), (128, 256, 1)])
def test_basic_block_residual_forward_backward_vs_float64(K, cin, cout, stride):
    """resnet.py:50-66 BasicBlock in train mode — conv+BN+ReLU, conv+BN, residual (identity, or 1x1 conv + BN downsample incl. the
    stride-2 variant of layer2.0), ReLU — forward, both gradient streams into the block input and every parameter gradient
    against float64 autograd: 1e-4.  Covers the `res` / `res_bn` / second-gradient-stream (`dz2`) paths of the fused passes."""
    import copy

    from istnet_b200 import image_engine as IE
    from istnet_b200.image import BasicBlock
    from istnet_b200.nhwc import Act

    torch.manual_seed(3)
    g = torch.Generator(device="cuda").manual_seed(17)
    down = None
    if stride != 1 or cin != cout:
        down = torch.nn.Sequential(torch.nn.Conv2d(cin, cout, 1, stride=stride, bias=False), torch.nn.BatchNorm2d(cout))
    blk = BasicBlock(cin, cout, stride=stride, downsample=down).cuda().train()
    sign = (torch.arange(cout, device="cuda") % 2 * 2 - 1).float()
    with torch.no_grad():
        for bn in [m for m in blk.modules() if isinstance(m, torch.nn.BatchNorm2d)]:
            torch.nn.init.uniform_(bn.weight, 0.5, 1.0)
            bn.momentum = 0.6
        blk.bn1.bias.copy_(3.0 * sign)
        blk.bn2.bias.copy_(4.0 * sign)
    B, H, W = 4, 16, 16
    x = torch.randn(B, cin, H, W, device="cuda", generator=g)
    b64 = copy.deepcopy(blk).double()
    x64 = x.double().requires_grad_(True)
    z64 = b64(x64)
    dz = torch.randn(z64.shape, device="cuda", dtype=torch.float64, generator=g)
    z64.backward(dz)
    u = _block_units(blk, "b")
    xin = Act(B, H, W, cin, x.permute(0, 2, 3, 1).contiguous())
    xin.pl = K.empty_planes(B, H, W, cin, "cuda")
    K.split(xin.f32, B * H * W, cin, xin.pl)
    tape = []
    out = IE._basic_block_fwd(u, "b", xin, True, True, tape)
    grads = {}
    d1, d2 = IE._basic_block_bwd(u, tape[0], dz.float().permute(0, 2, 3, 1).contiguous(), None, grads)
    K.join_side_streams()
    assert rel_err(out.f32, z64.detach().permute(0, 2, 3, 1)) < 1e-4
    assert rel_err(out.pl.float().sum(0)[..., :cout], z64.detach().permute(0, 2, 3, 1)) < 1e-4
    assert rel_err(d1 + d2, x64.grad.permute(0, 2, 3, 1)) < 1e-4, rel_err(d1 + d2, x64.grad.permute(0, 2, 3, 1))
    for (n, p), (_, p64) in zip(blk.named_parameters(), b64.named_parameters()):
        e = rel_err(grads[id(p)].reshape(p64.grad.shape), p64.grad)
        assert e < 1e-4, (n, e)
    for (n, b), (_, b_64) in zip(blk.named_buffers(), b64.named_buffers()):
        assert rel_err(b.double(), b_64.double()) < 1e-4, n


def test_stem_conv7x7_bn_relu_maxpool_forward_backward_vs_float64(K):
    """resnet.py:182-186: conv1 7x7/2 (im2col GEMM from the NCHW image) + BN(train) + ReLU + MaxPool(3,2,1), and the backward
    through max-pool / ReLU / BN into the 7x7 weight gradient (col2im is not needed: the image receives no gradient)."""
    import copy

    from istnet_b200 import image_engine as IE
    from istnet_b200.nhwc import ACT_RELU, ConvUnit

    torch.manual_seed(5)
    g = torch.Generator(device="cuda").manual_seed(19)
    conv = torch.nn.Conv2d(3, 64, 7, stride=2, padding=3, bias=False).cuda()
    bn = torch.nn.BatchNorm2d(64).cuda().train()
    with torch.no_grad():
        torch.nn.init.uniform_(bn.weight, 0.5, 1.0)
        bn.bias.copy_(0.5 * torch.randn(64, device="cuda", generator=g))
    B, H, W = 3, 64, 64
    rgb = torch.randn(B, 3, H, W, device="cuda", generator=g)
    c64, b64 = copy.deepcopy(conv).double(), copy.deepcopy(bn).double()
    z64 = F.max_pool2d(F.relu(b64(c64(rgb.double()))), 3, 2, 1)
    dz = torch.randn(z64.shape, device="cuda", dtype=torch.float64, generator=g)
    z64.backward(dz)
    u = {"conv1": ConvUnit(conv.weight, None, bn, ACT_RELU, k=7, stride=2, pad=3)}
    tape = []
    z = IE._stem_fwd(u, rgb, True, True, tape)
    grads = {}
    IE._stem_bwd(u, tape[0], dz.float().permute(0, 2, 3, 1).contiguous(), None, grads)
    K.join_side_streams()
    assert rel_err(z.f32, z64.detach().permute(0, 2, 3, 1)) < 1e-4
    assert rel_err(grads[id(conv.weight)], c64.weight.grad) < 1e-4, rel_err(grads[id(conv.weight)], c64.weight.grad)
    assert rel_err(grads[id(bn.weight)], b64.weight.grad) < 1e-4
    assert rel_err(grads[id(bn.bias)], b64.bias.grad) < 1e-4
    assert rel_err(bn.running_var.double(), b64.running_var) < 1e-4


def test_im2col_col2im_are_adjoint_and_match_unfold(K):
    """7x7/2 and 3x3/2 patch extraction (im2col_split) against F.unfold, and col2im against its autograd adjoint (fold)."""
    g = torch.Generator(device="cuda").manual_seed(23)
    for (C, k, stride, pad, H) in ((3, 7, 2, 3, 32), (64, 3, 2, 1, 16)):
        B = 2
        x = torch.randn(B, C, H, H, device="cuda", generator=g, dtype=torch.float64).requires_grad_(True)
        cols = F.unfold(x, k, padding=pad, stride=stride)  # (B, C*k*k, L), channel-major (c, r, s)
        Ho = (H + 2 * pad - k) // stride + 1
        ref = cols.view(B, C, k * k, Ho, Ho).permute(0, 3, 4, 2, 1).reshape(B, Ho, Ho, k * k * C)  # our K order: (r, s, c)
        a = K.im2col(x.detach().float().permute(0, 2, 3, 1).contiguous(), False, B, H, H, C, k, stride, pad)
        assert rel_err(a.pl.float().sum(0)[..., : k * k * C], ref.detach()) < 1e-6
        d = torch.randn(ref.shape, device="cuda", generator=g, dtype=torch.float64)
        ref.backward(d)
        dx = K.col2im(d.float().contiguous(), B, H, H, C, k, stride, pad)
        assert rel_err(dx, x.grad.permute(0, 2, 3, 1)) < 1e-6


def test_interp_rows_backward_vs_float64():
    """three_interpolate_grad on rows (interpolate_gpu.cu:121-148): d_feats[b, idx[b,j,q], :] += dout[b,j,:] * w[b,j,q]."""
    from istnet_b200 import rows_engine as RE

    g = torch.Generator(device="cuda").manual_seed(29)
    B, m, n, C = 3, 64, 200, 96
    feats = torch.randn(B, m, C, device="cuda", generator=g).requires_grad_(True)
    idx = torch.randint(0, m, (B, n, 3), device="cuda", generator=g, dtype=torch.int32)
    w = torch.rand(B, n, 3, device="cuda", generator=g)
    out = RE.interp_rows(feats, idx, w)
    cot = torch.randn(out.shape, device="cuda", generator=g)
    out.backward(cot)
    f64 = feats.detach().double().requires_grad_(True)
    ref = sum(torch.gather(f64, 1, idx[..., q].long()[..., None].expand(-1, -1, C)) * w.double()[..., q : q + 1] for q in range(3))
    ref.backward(cot.double())
    assert rel_err(out, ref) < 1e-6
    assert rel_err(feats.grad, f64.grad) < 1e-5


def test_pose_tail_pool_heads_ortho6d_vs_float64():
    """csrc/heads.cu: AdaptiveAvgPool1d(1) + 3 x [Linear, ReLU, Linear, ReLU, Linear] + Ortho6d2Mat (ist_net.py:228-264,
    utils/rotation_utils.py:4-28), forward and backward, against float64 autograd of the torch modules: 1e-5."""
    import copy

    from istnet_b200 import model as M

    torch.manual_seed(7)
    est = M.LightEstimator().cuda()
    B, N = 5, 96
    g = torch.Generator(device="cuda").manual_seed(31)
    feat = torch.randn(B * N, 512, device="cuda", generator=g).requires_grad_(True)
    r, t, s = M._PoseTailFn.apply(feat, B, N, *est._head_params())
    cot = [torch.randn(v.shape, device="cuda", generator=g) for v in (r, t, s)]
    torch.autograd.backward([r, t, s], cot)
    e64 = copy.deepcopy(est).double()
    f64 = feat.detach().double().requires_grad_(True)
    pooled = f64.view(B, N, -1).mean(1)
    r6 = e64.rotation_estimator(pooled)
    r_ref = M.ortho6d_to_mat(r6[:, :3].contiguous(), r6[:, 3:].contiguous())
    t_ref, s_ref = e64.translation_estimator(pooled), e64.size_estimator(pooled)
    torch.autograd.backward([r_ref, t_ref, s_ref], [c.double() for c in cot])
    for a, b in ((r, r_ref), (t, t_ref), (s, s_ref), (feat.grad, f64.grad)):
        assert rel_err(a, b) < 1e-5, rel_err(a, b)
    for name in ("rotation_estimator", "translation_estimator", "size_estimator"):
        for (n, p), (_, p64) in zip(getattr(est, name).named_parameters(), getattr(e64, name).named_parameters()):
            assert rel_err(p.grad, p64.grad) < 1e-5, (name, n, rel_err(p.grad, p64.grad))


def test_adam_flat_kernel_matches_torch_adam():
    """csrc/optim.cu against torch.optim.Adam (utils/solver.py:41-44) on identical gradients: 5 steps with a learning rate that
    changes every step through the device scalar (CyclicLR), weight decay, and the folded 1/world gradient scale."""
    import ctypes

    from istnet_b200 import _C
    from istnet_b200._C import c_float, c_ll, ptr

    g = torch.Generator().manual_seed(5)
    n = 4096 + 64
    for wd, scale in ((0.0, 1.0), (1e-2, 0.5)):
        p0 = torch.randn(n, generator=g)
        ref = torch.nn.Parameter(p0.clone().cuda())
        opt = torch.optim.Adam([ref], lr=1e-5, weight_decay=wd)
        p, m, v = p0.clone().cuda(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
        lr_dev = torch.zeros(1, device="cuda")
        step_dev = torch.zeros(1, dtype=torch.int64, device="cuda")
        for it, lr in enumerate((1e-5, 5.05e-4, 1e-3, 5.05e-4, 1e-5)):
            grad = (torch.randn(n, generator=g) * (10.0 ** float(torch.randint(-6, 1, (1,), generator=g)))).cuda()
            for gr in opt.param_groups:
                gr["lr"] = lr
            ref.grad = grad * scale
            opt.step()
            lr_dev.fill_(lr)
            _C.call("adam_flat", ptr(p), ptr(grad), ptr(m), ptr(v), c_ll(n), ptr(lr_dev), ctypes.c_double(0.9), ctypes.c_double(0.999), c_float(1e-8),
                    c_float(wd), c_float(scale), ptr(step_dev))
            _C.call("adam_tick", ptr(step_dev))
            assert int(step_dev.item()) == it + 1
            # same formula, FP32 both sides; torch evaluates the bias corrections in double on the host, the kernel in double on the device
            assert (p - ref.detach()).abs().max().item() <= 1e-4 * lr + 5e-7, (it, (p - ref.detach()).abs().max().item())
            st = opt.state[ref]
            assert rel_err(m, st["exp_avg"]) < 1e-6 and rel_err(v, st["exp_avg_sq"]) < 1e-6


def test_global_feature_as_per_instance_bias_equals_the_concatenated_form():
    """ist_net.py:172-173,257-258,325-326: conv([f | mean(f).expand]) evaluated as W_a f + (W_b mean(f) + b) with the bracket as a
    per-instance bias of the GEMM epilogue (conv_gemm bias_group), forward and backward, against float64 autograd of the concatenated
    form and against this repo's own concatenated path."""
    from istnet_b200 import model as M

    torch.manual_seed(9)
    b, n, c = 3, 256, 64
    seq = M._pt_mlp(2 * c, 96, 40).cuda()
    x = torch.randn(b * n, c, device="cuda")
    dout = torch.randn(b * n, 40, device="cuda")

    def run(fn):
        xi = x.clone().requires_grad_(True)
        for p in seq.parameters():
            p.grad = None
        out = fn(xi)
        out.backward(dout)
        return out.detach(), xi.grad.detach(), [p.grad.detach().clone() for p in seq.parameters()]

    assert M.GLOBAL_BIAS
    o1, dx1, g1 = run(lambda xi: M._mlp_global(seq, xi, b, n))
    o2, dx2, g2 = run(lambda xi: M._mlp(seq, M._with_global(xi, b, n)))
    seq64 = M._pt_mlp(2 * c, 96, 40).cuda().double()
    seq64.load_state_dict({k: v.double() for k, v in seq.state_dict().items()})
    x64 = x.double().requires_grad_(True)
    f = x64.view(b, n, c)
    cat = torch.cat([f, f.mean(1, keepdim=True).expand_as(f)], 2).transpose(1, 2)      # [b, 2c, n]
    o64 = seq64(cat).transpose(1, 2).reshape(b * n, -1)
    o64.backward(dout.double())
    g64 = [p.grad for p in seq64.parameters()]
    tol = 2e-5 if M.HEADS_NSPLIT >= 3 else 1e-4
    for name, mine, other, ref in (("out", o1, o2, o64), ("dx", dx1, dx2, x64.grad)):
        assert rel_err(mine, ref) < tol and rel_err(other, ref) < tol, (name, rel_err(mine, ref), rel_err(other, ref))
    for i, (a, c_, r) in enumerate(zip(g1, g2, g64)):
        assert rel_err(a, r.reshape(a.shape)) < tol and rel_err(c_, r.reshape(c_.shape)) < tol, (i, rel_err(a, r.reshape(a.shape)))
