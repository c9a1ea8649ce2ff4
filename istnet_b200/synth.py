"""Synthetic RGB-D instance crops for tests and bench.py (SURVEY.md §8(d), Appendix D input contract).

Points are *surface* samples (visible half of an ellipsoid in the camera frame) so that ball queries see a
realistic 2-D neighbour density: both the padded-ball and the saturated-ball branch of the ball query are hit
at every SA level.  Everything is generated on the CPU with a seeded torch.Generator, so the same seed gives
the same batch on every host.
"""
import math

import torch


def _random_rotations(b, g):
    q = torch.randn(b, 4, generator=g)
    q = q / q.norm(dim=1, keepdim=True)
    w, x, y, z = q.unbind(1)
    return torch.stack(
        (
            1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
            2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
            2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y),
        ),
        1,
    ).view(b, 3, 3)


def make_batch(batch, npts=1024, img=192, seed=1, duplicates=False, nclass=6, quantize=False):
    """Returns the dict the reference data pipeline hands to the model (provider/dataset.py:204-263)."""
    g = torch.Generator().manual_seed(int(seed))
    R = _random_rotations(batch, g)
    t = torch.empty(batch, 3)
    t[:, :2] = torch.rand(batch, 2, generator=g) * 0.6 - 0.3
    t[:, 2] = torch.rand(batch, generator=g) + 0.5
    axes = torch.rand(batch, 3, generator=g) * 0.11 + 0.04  # semi-axes, metres
    n_src = 300 if duplicates else npts
    # uniform directions, flipped onto the camera-facing hemisphere after rotation
    d = torch.randn(batch, n_src, 3, generator=g)
    d = d / d.norm(dim=2, keepdim=True)
    obj = d * axes[:, None, :]  # ellipsoid surface in the object frame
    cam = obj @ R.transpose(1, 2)
    flip = cam[..., 2:3] > 0  # keep the half facing the camera (towards -z)
    obj = torch.where(flip, -obj, obj)
    cam = obj @ R.transpose(1, 2)
    if duplicates:  # mirrors sampling with replacement at test time (dataset.py:387-392)
        pick = torch.randint(0, n_src, (batch, npts), generator=g)
        cam = torch.gather(cam, 1, pick[..., None].expand(-1, -1, 3))
    else:
        cam = cam + torch.clamp(0.001 * torch.randn(batch, npts, 3, generator=g), -0.005, 0.005)
    pts = (cam + t[:, None, :]).contiguous()
    if quantize:
        # 2^-12 m grid: every partial sum of <= 4096 coordinates is exact in FP32, so `pts - mean(pts)`
        # (ist_net.py:34-35) is bit-identical on every device / reduction order.
        pts = torch.round(pts * 4096.0) / 4096.0
    size = 2 * axes
    diag = size.norm(dim=1)
    qo = ((pts - t[:, None, :]) / diag[:, None, None]) @ R  # dataset.py:249
    return {
        "rgb": torch.randn(batch, 3, img, img, generator=g),
        "pts": pts.float().contiguous(),
        "choose": torch.randint(0, img * img, (batch, npts), generator=g, dtype=torch.int64),
        "category_label": torch.randint(0, nclass, (batch, 1), generator=g, dtype=torch.int64),
        "qo": qo.float().contiguous(),
        "rotation_label": R.contiguous(),
        "translation_label": t.contiguous(),
        "size_label": (size / diag[:, None]).contiguous(),
    }


def flops_per_instance(model="ist_net", npts=1024, train=True):
    """Algorithmic GFLOP per instance; used for the roofline line of bench.py."""
    table = {("ist_net", 1024): (42.05, 125.8), ("posenet_gt", 1024): (37.9, 43.8), ("ist_net", 4096): (59.5, 178.0)}
    fwd, both = table[(model, npts)]
    return both if train else fwd
