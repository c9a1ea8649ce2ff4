"""GPU parity of the nine point operators (through the C ABI) against the CPU oracle, bit-exact for every
index-producing op, plus — when oracle/_ref holds it — the reference's own CUDA extension compiled from
/root/reference (oracle/build_ref_ext.sh) as a third witness."""
import importlib.util
import os

import pytest
import torch

from istnet_b200.synth import make_batch
from oracle import pointops as po

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "pointnet2_ref", "_ext.so")


@pytest.fixture(scope="module")
def ext():
    from istnet_b200 import ext as e

    return e


@pytest.fixture(scope="module")
def ref_ext():
    if not os.path.exists(REF_SO):
        pytest.skip("reference extension not built (oracle/build_ref_ext.sh)")
    spec = importlib.util.spec_from_file_location("_ext", REF_SO)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def clouds(b, n, seed, dup=False):
    d = make_batch(b, n, 8, seed=seed, duplicates=dup, quantize=dup)
    return d["pts"] - d["pts"].mean(1, keepdim=True), d["qo"]


@pytest.mark.parametrize("n,m,dup", [(1024, 512, False), (512, 256, False), (256, 128, False), (128, 64, False),
                                      (1024, 512, True), (256, 512, False), (4096, 512, False), (300, 77, True), (1, 3, False), (33, 33, False)])
def test_fps_bit_exact(ext, n, m, dup):
    b = 3 if n < 4096 else 2
    for cloud in clouds(b, n, seed=n + m, dup=dup):
        want = po.furthest_point_sampling(cloud.contiguous(), m)
        got = ext.furthest_point_sampling(cloud.cuda().contiguous(), m).cpu()
        assert torch.equal(got, want)


def test_fps_tie_break_and_degenerate(ext):
    xyz = torch.zeros(1, 8, 3)
    xyz[0, 1:, 0] = 1.0
    assert ext.furthest_point_sampling(xyz.cuda(), 2).cpu()[0].tolist() == [0, 4]
    assert ext.furthest_point_sampling(torch.ones(2, 600, 3).cuda(), 9).cpu().tolist() == [[0] * 9] * 2
    assert ext.furthest_point_sampling(torch.randn(2, 16, 3).cuda(), 0).shape == (2, 0)


def test_fps_chain_equals_levelwise_reference_flow(ext):
    cloud, _ = clouds(4, 1024, seed=5)
    idxs, xyzs = ext.fps_chain(cloud.cuda().contiguous(), (512, 256, 128, 64))
    cur = cloud.contiguous()
    for l, m in enumerate((512, 256, 128, 64)):
        want = po.furthest_point_sampling(cur, m)
        nxt = po.gather_points(cur.transpose(1, 2).contiguous(), want).transpose(1, 2).contiguous()
        assert torch.equal(idxs[l].cpu(), want)
        assert torch.equal(xyzs[l].cpu(), nxt)
        cur = nxt


@pytest.mark.parametrize("n,m,r,ns", [(1024, 512, 0.01, 16), (1024, 512, 0.02, 32), (512, 256, 0.04, 32), (128, 64, 0.16, 32),
                                       (1024, 512, 0.05, 16), (4096, 512, 0.02, 32), (77, 13, 0.05, 5)])
def test_ball_query_bit_exact(ext, n, m, r, ns):
    for cloud in clouds(3, n, seed=n + ns, dup=(n == 512)):
        cloud = cloud.contiguous()
        cent = cloud[:, torch.randperm(n, generator=torch.Generator().manual_seed(1))[:m]].contiguous()
        want = po.ball_query(cent, cloud, r, ns)
        got = ext.ball_query(cent.cuda(), cloud.cuda(), r, ns).cpu()
        assert torch.equal(got, want)


def test_ball_query_edges(ext):
    xyz = torch.tensor([[[0.0, 0, 0], [0.5, 0, 0], [1.0, 0, 0], [0.25, 0, 0], [3.0, 0, 0]]])
    q = torch.tensor([[[0.0, 0, 0], [10.0, 0, 0], [0.5, 0, 0]]])
    idx = ext.ball_query(q.cuda(), xyz.cuda(), 0.5, 4).cpu()
    assert idx[0].tolist() == [[0, 3, 0, 0], [0, 0, 0, 0], [1, 3, 1, 1]]


@pytest.mark.parametrize("n,m", [(128, 64), (256, 128), (512, 256), (1024, 512), (50, 2)])
def test_three_nn_bit_exact(ext, n, m):
    cloud, _ = clouds(3, n, seed=n, dup=(n == 256))
    known = cloud[:, :m].contiguous()
    d2w, iw = po.three_nn(cloud.contiguous(), known)
    d2, i = ext.three_nn(cloud.cuda().contiguous(), known.cuda())
    assert torch.equal(i.cpu(), iw)
    assert torch.equal(d2.cpu(), d2w)


def test_gather_group_interpolate_and_grads(ext):
    g = torch.Generator().manual_seed(9)
    b, c, n, m, ns = 3, 67, 512, 256, 32
    feats = torch.randn(b, c, n, generator=g)
    idx = torch.randint(0, n, (b, m, ns), generator=g, dtype=torch.int32)
    i1 = torch.randint(0, n, (b, m), generator=g, dtype=torch.int32)
    i3 = torch.randint(0, n, (b, 700, 3), generator=g, dtype=torch.int32)
    w = torch.rand(b, 700, 3, generator=g)
    assert torch.equal(ext.group_points(feats.cuda(), idx.cuda()).cpu(), po.group_points(feats, idx))
    assert torch.equal(ext.gather_points(feats.cuda(), i1.cuda()).cpu(), po.gather_points(feats, i1))
    assert torch.equal(ext.three_interpolate(feats.cuda(), i3.cuda(), w.cuda()).cpu(), po.three_interpolate(feats, i3, w))
    go = torch.randn(b, c, m, ns, generator=g)
    assert torch.allclose(ext.group_points_grad(go.cuda(), idx.cuda(), n).cpu(), po.group_points_grad(go, idx, n), rtol=1e-5, atol=1e-5)
    go = torch.randn(b, c, m, generator=g)
    assert torch.allclose(ext.gather_points_grad(go.cuda(), i1.cuda(), n).cpu(), po.gather_points_grad(go, i1, n), rtol=1e-5, atol=1e-5)
    go = torch.randn(b, c, 700, generator=g)
    assert torch.allclose(ext.three_interpolate_grad(go.cuda(), i3.cuda(), w.cuda(), n).cpu(), po.three_interpolate_grad(go, i3, w, n), rtol=1e-5, atol=1e-5)


def test_full_size_properties(ext):
    """BASELINE cfg1/cfg4 sizes: properties that need no oracle."""
    for n in (1024, 4096):
        cloud = clouds(32, n, seed=n)[0].cuda().contiguous()
        idx = ext.furthest_point_sampling(cloud, 512)
        assert (idx[:, 0] == 0).all() and int(idx.min()) >= 0 and int(idx.max()) < n
        assert all(len(set(r.tolist())) == 512 for r in idx.cpu())  # distinct points => distinct samples
        cent = torch.gather(cloud, 1, idx.long()[..., None].expand(-1, -1, 3)).contiguous()
        bq = ext.ball_query(cent, cloud, 0.02, 32)
        # every row: strictly ascending prefix of hits, then padding with the first hit; centroid itself is a hit
        d = (torch.gather(cloud, 1, bq.long().reshape(32, -1, 1).expand(-1, -1, 3)).view(32, 512, 32, 3) - cent[:, :, None]).pow(2).sum(-1)
        assert (d < 0.02 * 0.02 * (1 + 1e-5)).all()
        inc = bq[:, :, 1:] > bq[:, :, :-1]
        pad = bq[:, :, 1:] == bq[:, :, :1]
        assert (inc | pad).all()


def test_reference_extension_agrees(ext, ref_ext):
    """The unmodified reference CUDA kernels (compiled for sm_100a) == C oracle == B200 kernels."""
    for n, m, dup in ((1024, 512, False), (512, 256, True), (256, 128, False), (300, 77, True)):
        for cloud in clouds(2, n, seed=3 * n, dup=dup):
            cg = cloud.cuda().contiguous()
            r_fps = ref_ext.furthest_point_sampling(cg, m)
            assert torch.equal(r_fps.cpu(), po.furthest_point_sampling(cloud.contiguous(), m))
            assert torch.equal(r_fps, ext.furthest_point_sampling(cg, m))
            cent = torch.gather(cg, 1, r_fps.long()[..., None].expand(-1, -1, 3)).contiguous()
            for r, ns in ((0.02, 16), (0.1, 32)):
                r_bq = ref_ext.ball_query(cent, cg, r, ns)
                assert torch.equal(r_bq.cpu(), po.ball_query(cent.cpu(), cloud.contiguous(), r, ns))
                assert torch.equal(r_bq, ext.ball_query(cent, cg, r, ns))
            d2r, ir = ref_ext.three_nn(cg, cent)
            d2, i = ext.three_nn(cg, cent)
            d2o, io = po.three_nn(cloud.contiguous(), cent.cpu())
            assert torch.equal(ir, i) and torch.equal(d2r, d2) and torch.equal(ir.cpu(), io) and torch.equal(d2r.cpu(), d2o)
            feats = torch.randn(2, 19, m, device="cuda")
            wgt = torch.rand(2, n, 3, device="cuda")
            assert torch.equal(ref_ext.three_interpolate(feats, ir, wgt), ext.three_interpolate(feats, i, wgt))
            assert torch.equal(ref_ext.three_interpolate(feats, ir, wgt).cpu(), po.three_interpolate(feats.cpu(), io, wgt.cpu()))
            assert torch.equal(ref_ext.group_points(feats, r_bq.clamp(max=m - 1)), ext.group_points(feats, r_bq.clamp(max=m - 1)))
