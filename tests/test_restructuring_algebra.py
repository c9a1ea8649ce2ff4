"""CPU, float64: the algebraic identities behind the three restructurings of the path hold against the
reference formulation evaluated with plain torch ops / autograd.  The CUDA kernels are tested against the same references on
the GPU (tests/test_gpu_kernels.py); these tests pin the mathematics itself and run anywhere."""
import torch
import torch.nn.functional as F

torch.set_default_dtype(torch.float32)


def test_psp_bottleneck_commutes_with_prior_upsampling():
    """modules.py:27-34: bottleneck(cat(up(stage_i(feats)) ..., feats)) == Wb_x feats + sum_i up(Wb_i stage_i(feats)) + b."""
    g = torch.Generator().manual_seed(0)
    B, C, H, W, Co, sizes = 2, 8, 12, 12, 16, (1, 2, 3, 6)
    x = torch.randn(B, C, H, W, generator=g, dtype=torch.float64)
    ws = [torch.randn(C, C, 1, 1, generator=g, dtype=torch.float64) for _ in sizes]
    wb = torch.randn(Co, C * (len(sizes) + 1), 1, 1, generator=g, dtype=torch.float64)
    b = torch.randn(Co, generator=g, dtype=torch.float64)
    priors = [F.interpolate(F.conv2d(F.adaptive_avg_pool2d(x, s), w), size=(H, W), mode="bilinear", align_corners=False) for s, w in zip(sizes, ws)]
    ref = F.relu(F.conv2d(torch.cat(priors + [x], 1), wb, b))
    acc = F.conv2d(x, wb[:, len(sizes) * C :], b)
    for i, (s, w) in enumerate(zip(sizes, ws)):
        t = F.conv2d(F.conv2d(F.adaptive_avg_pool2d(x, s), w), wb[:, i * C : (i + 1) * C])  # both 1x1 convolutions on the s x s map
        acc = acc + F.interpolate(t, size=(H, W), mode="bilinear", align_corners=False)
    assert torch.allclose(F.relu(acc), ref, atol=1e-12)


def test_head_batchnorm_from_input_moments_and_affine_backward():
    """modules.py:64-66 + ist_net.py:42-45: train-mode BN statistics of y = W x + b from sum(x) and X^T X; BN backward of a gradient
    that is non-zero only at gathered pixels = sparse rows + a term affine in x (image_engine._head_forward/_head_backward)."""
    g = torch.Generator().manual_seed(1)
    P, C, Co, R, eps = 500, 6, 10, 40, 1e-5
    x = torch.randn(P, C, generator=g, dtype=torch.float64) + 0.3
    W = torch.randn(Co, C, generator=g, dtype=torch.float64).requires_grad_(True)
    b = torch.randn(Co, generator=g, dtype=torch.float64).requires_grad_(True)
    gamma = (torch.rand(Co, generator=g, dtype=torch.float64) + 0.5).requires_grad_(True)
    beta = torch.randn(Co, generator=g, dtype=torch.float64).requires_grad_(True)
    xr = x.clone().requires_grad_(True)
    sel = torch.randint(0, P, (R,), generator=g)
    sel[1] = sel[0]  # a pixel gathered twice
    cot = torch.randn(R, Co, generator=g, dtype=torch.float64)
    # reference: dense conv -> BN(train) -> gather
    y = xr @ W.t() + b
    mu, var = y.mean(0), y.var(0, unbiased=False)
    u = (y - mu) / torch.sqrt(var + eps) * gamma + beta
    u[sel].backward(cot)
    # statistics from the moments of x
    sx, S = x.sum(0), x.t() @ x
    mx = sx / P
    mean = W.detach() @ mx + b.detach()
    varm = ((W.detach() @ (S / P - torch.outer(mx, mx))) * W.detach()).sum(1)
    assert torch.allclose(mean, mu.detach(), atol=1e-12) and torch.allclose(varm, var.detach(), atol=1e-12)
    # backward: sparse rows + affine term
    inv = 1.0 / torch.sqrt(varm + eps)
    Wd, bd, gd = W.detach(), b.detach(), gamma.detach()
    yg = x[sel] @ Wd.t() + bd
    xhat = (yg - mean) * inv
    gs = cot  # no activation here: g = cotangent on the gathered rows
    sg, sgx = gs.sum(0), (gs * xhat).sum(0)
    assert torch.allclose(sgx, gamma.grad, atol=1e-10) and torch.allclose(sg, beta.grad, atol=1e-10)
    a, d = gd * inv * sg / P, gd * inv * inv * sgx / P
    dys = gs * (gd * inv)
    dW = dys.t() @ x[sel] - torch.outer(a, sx) - d[:, None] * (Wd @ S + torch.outer(bd - mean, sx))
    assert torch.allclose(dW, W.grad, atol=1e-9)
    assert float(b.grad.abs().max()) < 1e-10  # a bias feeding a train-mode BatchNorm has zero gradient
    A, c = Wd.t() @ (d[:, None] * Wd), Wd.t() @ (a + d * (bd - mean))
    dx = -(x @ A.t()) - c
    dx.index_add_(0, sel, dys @ Wd)
    assert torch.allclose(dx, xr.grad, atol=1e-9)


def test_sa_layer0_on_points_equals_layer0_on_grouped_rows():
    """pointnet2_utils.py:335-367 + the first SharedMLP conv: W [xyz_j - c_i ; f_j] == Wx (xyz_j - c_i) + (F Wf^T)[j], and the
    backward: dF = scatter(dy0) Wf, dWf = scatter(dy0)^T F, dWx = sum_rows dy0 (x) (xyz_j - c_i)  (rows_engine._sa_l0_*)."""
    g = torch.Generator().manual_seed(2)
    N, M, ns, C, C0 = 30, 7, 5, 4, 6
    xyz = torch.randn(N, 3, generator=g, dtype=torch.float64)
    cent = xyz[torch.randperm(N, generator=g)[:M]]
    idx = torch.randint(0, N, (M, ns), generator=g)
    feats = torch.randn(N, C, generator=g, dtype=torch.float64).requires_grad_(True)
    W = torch.randn(C0, 3 + C, generator=g, dtype=torch.float64).requires_grad_(True)
    grouped = torch.cat([xyz[idx] - cent[:, None], feats[idx]], -1)  # (M, ns, 3+C)
    y_ref = grouped @ W.t()
    cot = torch.randn(M, ns, C0, generator=g, dtype=torch.float64)
    y_ref.backward(cot)
    Wd, Fd = W.detach(), feats.detach()
    u = Fd @ Wd[:, 3:].t()
    rel = xyz[idx] - cent[:, None]
    y = u[idx] + rel @ Wd[:, :3].t()
    assert torch.allclose(y, y_ref.detach(), atol=1e-12)
    dU = torch.zeros(N, C0, dtype=torch.float64).index_add_(0, idx.reshape(-1), cot.reshape(-1, C0))
    assert torch.allclose(dU @ Wd[:, 3:], feats.grad, atol=1e-12)
    dW = torch.cat([cot.reshape(-1, C0).t() @ rel.reshape(-1, 3), dU.t() @ Fd], 1)
    assert torch.allclose(dW, W.grad, atol=1e-12)


def test_operand_plane_products_reach_the_documented_accuracy():
    """x = p0 + p1 (+ p2) in bf16 planes, products of all plane pairs with i + j < n: ~3e-6 (n = 2) and below FP32
    rounding (n = 3) per product, against float64."""
    g = torch.Generator().manual_seed(3)
    a = torch.randn(4096, generator=g) * torch.logspace(-3, 3, 4096)
    b = torch.randn(4096, generator=g)

    def planes(x, n):
        out, r = [], x.clone()
        for _ in range(n):
            h = r.to(torch.bfloat16).float()
            out.append(h)
            r = r - h
        return out

    exact = a.double() * b.double()
    for n, bound in ((2, 2.0 ** -15), (3, 2.0 ** -22)):
        pa, pb = planes(a, n), planes(b, n)
        prod = sum(pa[i].double() * pb[j].double() for i in range(n) for j in range(n) if i + j < n)
        rel = ((prod - exact).abs() / exact.abs().clamp_min(1e-300)).max().item()
        assert rel < bound, (n, rel)
    p3 = planes(a, 3)
    assert ((p3[0] + p3[1] + p3[2]).double() - a.double()).abs().max().item() <= (a.abs().max().item() * 2.0 ** -24)
