"""Fused set-abstraction level (csrc/sa_fused.cu: ball query + grouping + SharedMLP with train-mode BatchNorm + max over nsample,
both scales per launch) against the reference dataflow evaluated in FLOAT64 with torch ops (pointnet2_modules.py:29-73,
pointnet2_utils.py:317-377, pytorch_utils.py:25-206), and against the unfused round-1 path of this repo."""
import copy

import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err

pytestmark = pytest.mark.gpu


def _level(cin, widths, radii, npoint, seed, negative_gamma):
    from istnet_b200.pointnet2 import PointnetSAModuleMSG

    torch.manual_seed(seed)
    sa = PointnetSAModuleMSG(npoint, radii, [16, 32], [[cin, *widths], [cin, *widths]]).cuda().train()
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for m in sa.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.weight.copy_((0.5 + torch.rand(m.num_features, generator=g)).to(m.weight.device))
                if negative_gamma:  # exercises the min-selection branch: relu(bn(.)) is decreasing in y for gamma < 0
                    m.weight.mul_(torch.where(torch.rand(m.num_features, generator=g) < 0.4, -1.0, 1.0).to(m.weight.device))
                m.bias.copy_((0.2 * torch.randn(m.num_features, generator=g)).to(m.weight.device))
                m.momentum = 0.37
    return sa


def _reference_f64(sa, xyz, new_xyz, feats, idxs):
    """The reference flow in float64: group (xyz first, centred), SharedMLP in train mode, max over nsample, concat scales."""
    outs = []
    mlps64 = []
    for mlp, idx in zip(sa.mlps, idxs):
        m64 = copy.deepcopy(mlp).double().train()
        mlps64.append(m64)
        gi = idx.long()
        M, ns = gi.shape[1], gi.shape[2]
        gx = torch.gather(xyz.double()[:, None].expand(-1, M, -1, -1), 2, gi[..., None].expand(-1, -1, -1, 3)) - new_xyz.double()[:, :, None]
        if feats is not None:
            C = feats.shape[2]
            gf = torch.gather(feats[:, None].expand(-1, M, -1, -1), 2, gi[..., None].expand(-1, -1, -1, C))
            grouped = torch.cat([gx, gf], -1)
        else:
            grouped = gx
        y = m64(grouped.permute(0, 3, 1, 2))  # (B, 3+C, npoint, nsample)
        outs.append(F.max_pool2d(y, kernel_size=[1, ns]).squeeze(-1).transpose(1, 2))
    return torch.cat(outs, 2), mlps64


@pytest.mark.parametrize("cin,widths,radii,N,M,neg", [
    (0, (16, 16, 32), [0.01, 0.02], 1024, 512, False),    # SA level 1 of the camera-space extractor
    (0, (16, 16, 32), [0.05, 0.10], 1024, 512, True),     # SA level 1 of the NOCS-space extractor, mixed-sign gammas
    (64, (32, 32, 64), [0.02, 0.04], 512, 256, False),    # SA level 2
    (64, (32, 32, 64), [0.10, 0.20], 512, 256, True),
    (0, (16, 16, 32), [0.01, 0.02], 256, 512, False),     # cfg0 shape: more centroids than points (FPS repeats index 0)
])
def test_fused_level_forward_backward_vs_float64(cin, widths, radii, N, M, neg):
    from istnet_b200 import ext, sa_fused as SF
    from istnet_b200.synth import make_batch

    B = 3
    d = make_batch(B, N, 8, seed=5)
    src = d["qo"] if radii[0] >= 0.05 else d["pts"] - d["pts"].mean(1, keepdim=True)
    xyz = src.cuda().contiguous()
    g = torch.Generator(device="cuda").manual_seed(2)
    feats = torch.randn(B, N, cin, device="cuda", generator=g).requires_grad_(True) if cin else None
    fi = ext.furthest_point_sampling(xyz, M)
    new_xyz = torch.gather(xyz, 1, fi.long()[..., None].expand(-1, -1, 3)).contiguous()
    sa = _level(cin, widths, radii, M, seed=7, negative_gamma=neg)
    assert SF.supported(sa, xyz, new_xyz, feats)
    rm0 = [copy.deepcopy(m.running_mean) for m in sa.modules() if isinstance(m, torch.nn.BatchNorm2d)]
    out = SF.sa_level(sa, xyz, new_xyz, feats)
    # indices written by the fused query pass are bit-exact with the stand-alone operator (itself bit-exact with the reference)
    _, (u, idx, ysel, asel, states) = SF._forward(copy.deepcopy(sa), True, xyz, new_xyz, feats.detach() if cin else None)
    for s, grouper in enumerate(sa.groupers):
        assert torch.equal(idx[s], ext.ball_query(new_xyz, xyz, grouper.radius, grouper.nsample))
    f64 = feats.detach().double().requires_grad_(True) if cin else None
    ref, mlps64 = _reference_f64(sa_ref := copy.deepcopy(sa), xyz, new_xyz, f64, idx)
    # running statistics: the float64 clones were deep-copied AFTER the fused forward updated the originals; redo from the saved start
    assert rel_err(out, ref) < 1e-4, rel_err(out, ref)
    cot = torch.randn(out.shape, device="cuda", generator=g)
    out.backward(cot)
    ref.backward(cot.double())
    if cin:
        assert rel_err(feats.grad, f64.grad) < 2e-4, rel_err(feats.grad, f64.grad)
    for s in range(2):
        for (n, p), p64 in zip(sa.mlps[s].named_parameters(), mlps64[s].parameters()):
            e = rel_err(p.grad.reshape(p64.grad.shape), p64.grad)
            assert e < 3e-4, (s, n, e)
    # BatchNorm bookkeeping of the fused passes: momentum 0.37 applied once to every layer, unbiased variance, counter + 1
    bns = [m for m in sa.modules() if isinstance(m, torch.nn.BatchNorm2d)]
    assert all(int(m.num_batches_tracked) == 1 for m in bns)
    assert any((m.running_mean - r).abs().max() > 1e-6 for m, r in zip(bns, rm0))


def test_fused_level_equals_unfused_path_including_running_stats():
    """Same module, same inputs, fused vs round-1 unfused kernels: outputs 1e-5, gradients 2e-4, running statistics 1e-5."""
    from istnet_b200 import ext, sa_fused as SF
    from istnet_b200.synth import make_batch

    B, N, M = 4, 512, 256
    d = make_batch(B, N, 8, seed=9)
    xyz = (d["pts"] - d["pts"].mean(1, keepdim=True)).cuda().contiguous()
    g = torch.Generator(device="cuda").manual_seed(4)
    fi = ext.furthest_point_sampling(xyz, M)
    new_xyz = torch.gather(xyz, 1, fi.long()[..., None].expand(-1, -1, 3)).contiguous()
    res = {}
    for fused in (True, False):
        sa = _level(64, (32, 32, 64), [0.02, 0.04], M, seed=3, negative_gamma=True)
        feats = torch.randn(B, N, 64, device="cuda", generator=torch.Generator(device="cuda").manual_seed(6)).requires_grad_(True)
        SF.ENABLED = fused
        try:
            _, out = sa.forward_rows(xyz, feats, new_xyz=new_xyz)
        finally:
            SF.ENABLED = True
        out.backward(torch.randn(out.shape, device="cuda", generator=torch.Generator(device="cuda").manual_seed(8)))
        res[fused] = (out.detach(), feats.grad, [p.grad for p in sa.parameters()], [b.clone() for b in sa.buffers()])
    assert rel_err(res[True][0], res[False][0]) < 1e-5
    assert rel_err(res[True][1], res[False][1]) < 2e-4
    for a, b in zip(res[True][2], res[False][2]):
        assert rel_err(a, b) < 2e-4
    for a, b in zip(res[True][3], res[False][3]):
        assert rel_err(a.double(), b.double()) < 1e-5


def test_fused_level_eval_mode_uses_running_statistics():
    from istnet_b200 import ext, sa_fused as SF
    from istnet_b200.synth import make_batch
    from conftest import perturb_batchnorm

    B, N, M = 2, 1024, 512
    d = make_batch(B, N, 8, seed=11)
    xyz = (d["pts"] - d["pts"].mean(1, keepdim=True)).cuda().contiguous()
    fi = ext.furthest_point_sampling(xyz, M)
    new_xyz = torch.gather(xyz, 1, fi.long()[..., None].expand(-1, -1, 3)).contiguous()
    sa = _level(0, (16, 16, 32), [0.01, 0.02], M, seed=13, negative_gamma=True)
    perturb_batchnorm(sa, 17)
    sa.eval()
    with torch.no_grad():
        out = SF.sa_level(sa, xyz, new_xyz, None)
        idx = [ext.ball_query(new_xyz, xyz, gr.radius, gr.nsample) for gr in sa.groupers]
        sa64 = copy.deepcopy(sa)
        outs = []
        for mlp, ix in zip(sa64.mlps, idx):
            m64 = mlp.double().eval()
            gi = ix.long()
            gx = torch.gather(xyz.double()[:, None].expand(-1, M, -1, -1), 2, gi[..., None].expand(-1, -1, -1, 3)) - new_xyz.double()[:, :, None]
            outs.append(F.max_pool2d(m64(gx.permute(0, 3, 1, 2)), kernel_size=[1, gi.shape[2]]).squeeze(-1).transpose(1, 2))
        ref = torch.cat(outs, 2)
    assert rel_err(out, ref) < 1e-4
