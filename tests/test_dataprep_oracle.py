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


AUG = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dataprep_aug.npz")


def test_label_and_augmentation_restatements_match_the_reference_goldens():
    """NOCS labels (dataset.py:236-250) and the two default augmentations (data_augmentation.py:45-130): the goldens come from the
    reference's own expressions / functions (tests/gen_golden_dataprep.py); oracle and the product's host-side label logic against them."""
    import torch

    from istnet_b200.dataprep import augment_label_params, canonical_labels

    g, a = np.load(GOLD), np.load(AUG)
    B = len(a["do_bb"])
    rot64, par = canonical_labels(a["rotation"], a["translation"], a["size"], a["symmetric"])
    for b in range(B):
        pts64 = g["pts"][b].astype(np.float32) + g["noise"][b]
        qo, rot = R.nocs_labels(pts64, a["rotation"][b], a["translation"][b], a["size"][b], bool(a["symmetric"][b]))
        assert np.array_equal(qo, a["qo"][b]), b
        assert np.array_equal(np.asarray(rot, dtype=np.float32), a["rotation_label"][b]), b
        assert np.array_equal(rot64[b].astype(np.float32), a["rotation_label"][b]) and np.array_equal(par[b, 4:], np.asarray(rot, np.float64).reshape(9))
        out = R.augment_points(g["pts_jitter"][b], a["rotation_label"][b], a["translation"][b], a["size"][b], a["qo"][b], int(a["sym0"][b]),
                               bool(a["do_bb"][b]), a["aug_bb"][b], bool(a["do_rt"][b]), a["aug_t"][b], a["aug_R"][b])
        for got, key in zip(out, ("out_pts", "out_R", "out_t", "out_s", "out_qo")):
            assert np.allclose(got, a[key][b], rtol=2e-6, atol=2e-7), (b, key, np.abs(got - a[key][b]).max())
    _, Rl, tl, sl = augment_label_params(a["rotation_label"], a["translation"], a["size"], a["sym0"], a["do_bb"], a["aug_bb"], a["do_rt"],
                                         a["aug_t"], a["aug_R"])
    assert torch.allclose(Rl, torch.from_numpy(a["out_R"]), rtol=2e-6, atol=2e-7)
    assert torch.allclose(tl, torch.from_numpy(a["out_t"]), rtol=2e-6, atol=2e-7) and torch.allclose(sl, torch.from_numpy(a["out_s"]), rtol=2e-6, atol=2e-7)
