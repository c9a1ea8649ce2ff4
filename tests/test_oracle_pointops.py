"""CPU tests of the oracle's C restatement of the nine point ops (oracle/pointops_ref.c) against a literal,
thread-by-thread Python simulation of the reference kernels and against the edge cases SURVEY.md §8c lists."""
import math

import numpy as np
import pytest
import torch

from oracle import pointops as po

f32 = np.float32


def _sq(dx, dy, dz):
    # fmaf(dz,dz,fmaf(dx,dx,dy*dy)) (the reference's SASS order) evaluated exactly: products/sums in float64 are exact
    # for float32 inputs up to the single rounding of each fused step.
    t = f32(f32(dy) * f32(dy))
    t = f32(np.float64(dx) * np.float64(dx) + np.float64(t))
    return f32(np.float64(dz) * np.float64(dz) + np.float64(t))


def sim_fps(xyz, m):
    """Literal simulation of furthest_point_sampling_kernel (sampling_gpu.cu:74-178), one instance."""
    n = xyz.shape[0]
    S = max(min(1 << int(math.log(n) / math.log(2.0)), 512), 1)
    temp = np.full(n, 1e10, f32)
    out = [0]
    old = 0
    for _ in range(1, m):
        dists, dists_i = np.zeros(S, f32), np.zeros(S, np.int64)
        for t in range(S):
            best, besti = f32(-1), 0
            for k in range(t, n, S):
                d = _sq(xyz[k, 0] - xyz[old, 0], xyz[k, 1] - xyz[old, 1], xyz[k, 2] - xyz[old, 2])
                d2 = min(d, temp[k])
                temp[k] = d2
                if d2 > best:
                    best, besti = d2, k
            dists[t], dists_i[t] = best, besti
        s = S // 2
        while s >= 1:
            for t in range(s):
                if dists[t + s] > dists[t]:
                    dists_i[t] = dists_i[t + s]
                dists[t] = max(dists[t], dists[t + s])
            s //= 2
        old = int(dists_i[0])
        out.append(old)
    return np.array(out, np.int32)


@pytest.mark.parametrize("n,m,dup", [(64, 16, False), (100, 40, False), (96, 48, True), (32, 48, False)])
def test_fps_matches_thread_simulation(n, m, dup):
    g = torch.Generator().manual_seed(n * 7 + m)
    xyz = torch.randn(2, n, 3, generator=g)
    if dup:  # exact duplicates => ties decided by the block-reduction order
        xyz = xyz[:, torch.randint(0, n // 3, (n,), generator=g)]
    got = po.furthest_point_sampling(xyz.contiguous(), m)
    for b in range(2):
        np.testing.assert_array_equal(got[b].numpy(), sim_fps(xyz[b].numpy(), m))


def test_fps_all_points_identical_and_npoint_gt_n():
    xyz = torch.ones(1, 16, 3)
    assert po.furthest_point_sampling(xyz, 8).tolist() == [[0] * 8]
    # npoint > N (cfg0: 512 samples from 256 points): once every min-distance is 0 index 0 repeats
    xyz = torch.randn(1, 8, 3, generator=torch.Generator().manual_seed(0))
    idx = po.furthest_point_sampling(xyz, 12)[0].tolist()
    assert sorted(idx[:8]) == list(range(8)) and idx[8:] == [0, 0, 0, 0]


def test_fps_tie_break_bit_reversed_thread():
    # n=8 => S=8 threads. Points 1..7 all at the same distance from point 0: the tree (s=4,2,1) keeps the left
    # operand on ties, so the winner is the candidate whose 3-bit id has the smallest bit reversal: id 4 (100b -> 001b).
    xyz = torch.zeros(1, 8, 3)
    xyz[0, 1:, 0] = 1.0
    assert po.furthest_point_sampling(xyz, 2)[0].tolist() == [0, 4]


def test_ball_query_edges():
    xyz = torch.tensor([[[0.0, 0, 0], [0.5, 0, 0], [1.0, 0, 0], [0.25, 0, 0], [3.0, 0, 0]]])
    q = torch.tensor([[[0.0, 0, 0], [10.0, 0, 0], [0.5, 0, 0]]])
    idx = po.ball_query(q, xyz, 0.5, 4)
    # centroid 0: d2 == r2 for point 1 is REJECTED (strict <); hits 0,3 then padded with the first hit
    assert idx[0, 0].tolist() == [0, 3, 0, 0]
    assert idx[0, 1].tolist() == [0, 0, 0, 0]  # empty ball => zeros
    assert idx[0, 2].tolist() == [1, 3, 1, 1]
    idx = po.ball_query(q, xyz, 100.0, 3)  # saturated ball: first nsample in index order
    assert idx[0, 0].tolist() == [0, 1, 2]


def test_three_nn_ties_and_order():
    known = torch.tensor([[[1.0, 0, 0], [-1.0, 0, 0], [0, 1.0, 0], [0, 0, 2.0]]])
    unknown = torch.zeros(1, 1, 3)
    d2, idx = po.three_nn(unknown, known)
    assert idx[0, 0].tolist() == [0, 1, 2] and d2[0, 0].tolist() == [1.0, 1.0, 1.0]  # ties keep the lower k
    d2, idx = po.three_nn(unknown, known[:, :2])  # fewer than 3 known points: sentinel 1e40 -> inf, index 0
    assert idx[0, 0].tolist() == [0, 1, 0] and math.isinf(d2[0, 0, 2].item())


def test_group_gather_interpolate_roundtrip():
    g = torch.Generator().manual_seed(3)
    feats = torch.randn(2, 5, 20, generator=g)
    idx = torch.randint(0, 20, (2, 7, 4), generator=g, dtype=torch.int32)
    out = po.group_points(feats, idx)
    ref = torch.gather(feats[:, :, None, :].expand(-1, -1, 7, -1), 3, idx.long()[:, None].expand(-1, 5, -1, -1))
    assert torch.equal(out, ref)
    go = torch.randn_like(out)
    gref = torch.zeros_like(feats).scatter_add_(2, idx.long().reshape(2, 1, -1).expand(-1, 5, -1), go.reshape(2, 5, -1))
    assert torch.allclose(po.group_points_grad(go, idx, 20), gref, atol=1e-6)
    i1 = torch.randint(0, 20, (2, 9), generator=g, dtype=torch.int32)
    assert torch.equal(po.gather_points(feats, i1), torch.gather(feats, 2, i1.long()[:, None].expand(-1, 5, -1)))
    w = torch.rand(2, 9, 3, generator=g)
    i3 = torch.randint(0, 20, (2, 9, 3), generator=g, dtype=torch.int32)
    o = po.three_interpolate(feats, i3, w)
    ref = sum(torch.gather(feats, 2, i3[..., q].long()[:, None].expand(-1, 5, -1)) * w[..., q][:, None] for q in range(3))
    assert torch.allclose(o, ref, atol=1e-6)
    go = torch.randn_like(o)
    gi = po.three_interpolate_grad(go, i3, w, 20)
    gref = torch.zeros_like(feats)
    for q in range(3):
        gref.scatter_add_(2, i3[..., q].long()[:, None].expand(-1, 5, -1), go * w[..., q][:, None])
    assert torch.allclose(gi, gref, atol=1e-5)
    assert torch.allclose(po.gather_points_grad(go, i1, 20), torch.zeros_like(feats).scatter_add_(2, i1.long()[:, None].expand(-1, 5, -1), go), atol=1e-6)
