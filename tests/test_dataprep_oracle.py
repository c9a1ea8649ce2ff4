"""SURVEY.md §8f row f3 — the CPU restatement of the per-instance input preparation (oracle/dataprep_ref.py) against golden vectors
produced by cv2 / torchvision / numpy with the reference's expressions (tests/gen_golden_dataprep.py), and against the live
libraries where they are importable."""
import os

import numpy as np
import pytest

from oracle import dataprep_ref as R

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dataprep.npz")


def test_oracle_reproduces_the_library_goldens_bit_exactly():
    g = np.load(GOLD)
    S, norm = int(g["S"]), float(g["norm_scale"])
    for i, box in enumerate(g["boxes"]):
        f, win = int(box[0]), tuple(int(v) for v in box[1:])
        rgb, pts, cho = R.prepare_instance(g["frames"][f], g["depth"][f], win, g["choose_in"][i].astype(np.int64), tuple(g["intrinsics"]), norm, S)
        assert np.array_equal(rgb, g["rgb"][i]), i          # cv2.resize (8-bit bilinear) + ToTensor + Normalize
        assert np.array_equal(pts, g["pts"][i]), i          # back-projection
        assert np.array_equal(cho, g["choose"][i]), i       # choose re-mapping
        _, ptsn, _ = R.prepare_instance(g["frames"][f], g["depth"][f], win, g["choose_in"][i].astype(np.int64), tuple(g["intrinsics"]), norm, S,
                                        noise=g["noise"][i])
        assert np.array_equal(ptsn, g["pts_jitter"][i]), i  # + float64 jitter, one rounding


def test_resize_restatement_equals_live_cv2_on_random_crops():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(3)
    for t in range(40):
        h = int(rng.integers(4, 441))
        w = h if t % 3 else int(rng.integers(4, 441))
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        for S in (192, 64):
            assert np.array_equal(R.resize_linear_u8(img, S), cv2.resize(img, (S, S), interpolation=cv2.INTER_LINEAR)), (h, w, S)


def test_normalize_restatement_equals_live_torchvision():
    T = pytest.importorskip("torchvision.transforms")
    rng = np.random.default_rng(4)
    tr = T.Compose([T.ToTensor(), T.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
    x = rng.integers(0, 256, (192, 192, 3), dtype=np.uint8)
    assert np.array_equal(tr(x).numpy(), R.normalize_u8(x))


def test_get_bbox_and_choose_sampler_host_logic():
    import torch

    from istnet_b200.dataprep import get_bbox, sample_choose

    assert get_bbox((100, 200, 180, 300)) == (80, 200, 190, 310)      # 100 px -> window 120, centred
    assert get_bbox((0, 0, 30, 30)) == (0, 40, 0, 40)                  # shifted back inside the image
    assert get_bbox((400, 600, 480, 640)) == (360, 480, 520, 640)    # pushed back from the bottom-right corner
    assert get_bbox((0, 0, 480, 640))[1] - get_bbox((0, 0, 480, 640))[0] == 440  # capped at 440
    # sampler (CPU tensors are fine for this torch-only helper)
    mask = torch.zeros(1, 40, 40, dtype=torch.bool)
    mask[0, 5:9, 5:9] = True
    g = torch.Generator().manual_seed(1)
    ch, ok = sample_choose(mask, [(0, 0, 40, 0, 40), (0, 20, 40, 20, 40)], 32, generator=g)
    assert bool(ok[0]) and not bool(ok[1])
    valid = set((mask[0].reshape(-1).nonzero().squeeze(1)).tolist())
    assert set(ch[0].tolist()) <= valid and len(ch[0]) == 32        # 16 valid pixels < 32: drawn with replacement
    ch2, _ = sample_choose(mask, [(0, 0, 40, 0, 40)], 8, generator=g)
    assert len(set(ch2[0].tolist())) == 8 and set(ch2[0].tolist()) <= valid  # more valid pixels than samples: without replacement
