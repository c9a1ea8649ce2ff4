"""SURVEY.md §8f row f3 on the GPU: csrc/dataprep.cu through the C ABI against (a) the golden vectors produced by cv2 / torchvision /
numpy with the reference's expressions (tests/golden/dataprep.npz) and (b) the CPU restatement at the reference's full sizes
(480x640 frames, 192x192 crops, 1024 points).  Everything is integer / correctly-rounded arithmetic: the bar is bit-exact."""
import os

import numpy as np
import pytest
import torch

from istnet_b200 import dataprep as D
from oracle import dataprep_ref as R

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dataprep.npz")


def _run(frames, depth, boxes, choose, intr, S, norm, noise=None):
    out = D.prepare_instances(torch.from_numpy(frames).cuda(), torch.from_numpy(depth).cuda(), torch.from_numpy(boxes).cuda(),
                              torch.from_numpy(choose).cuda(), intr, img_size=S, norm_scale=norm,
                              noise=None if noise is None else torch.from_numpy(noise))
    return {k: v.cpu().numpy() for k, v in out.items()}


def test_matches_the_library_goldens_bit_exactly():
    g = np.load(GOLD)
    S, norm = int(g["S"]), float(g["norm_scale"])
    out = _run(g["frames"], g["depth"], g["boxes"], g["choose_in"], tuple(g["intrinsics"]), S, norm)
    assert np.array_equal(out["rgb"], g["rgb"])
    assert np.array_equal(out["pts"], g["pts"])
    assert np.array_equal(out["choose"], g["choose"]) and out["choose"].dtype == np.int64
    outn = _run(g["frames"], g["depth"], g["boxes"], g["choose_in"], tuple(g["intrinsics"]), S, norm, noise=g["noise"])
    assert np.array_equal(outn["pts"], g["pts_jitter"])


def test_full_size_frames_match_the_cpu_restatement():
    rng = np.random.default_rng(11)
    F, H, W, S, N = 3, 480, 640, 192, 1024
    frames = rng.integers(0, 256, (F, H, W, 3), dtype=np.uint8)
    depth = rng.uniform(300, 2500, (F, H, W)).astype(np.float32)
    dets = [(0, (100, 200, 180, 300)), (0, (0, 0, 30, 30)), (1, (400, 600, 480, 640)), (1, (0, 0, 480, 640)), (2, (17, 333, 250, 401)),
            (2, (230, 10, 300, 200)), (0, (50, 50, 61, 58))]
    boxes = np.array([(f,) + D.get_bbox(b) for f, b in dets], dtype=np.int32)
    choose = np.stack([rng.integers(0, (b[2] - b[1]) * (b[4] - b[3]), N) for b in boxes]).astype(np.int32)
    intr = (591.0125, 590.16775, 322.525, 244.11084)  # provider/dataset.py:39
    out = _run(frames, depth, boxes, choose, intr, S, 1000.0)
    for i, b in enumerate(boxes):
        rgb, pts, cho = R.prepare_instance(frames[b[0]], depth[b[0]], tuple(int(v) for v in b[1:]), choose[i].astype(np.int64), intr, 1000.0, S)
        assert np.array_equal(out["rgb"][i], rgb), i
        assert np.array_equal(out["pts"][i], pts), i
        assert np.array_equal(out["choose"][i], cho), i
        assert out["choose"][i].min() >= 0 and out["choose"][i].max() < S * S


def test_prepared_batch_feeds_the_model_and_cpu_tensors_are_refused():
    from istnet_b200 import model as M

    rng = np.random.default_rng(12)
    frames = rng.integers(0, 256, (1, 480, 640, 3), dtype=np.uint8)
    depth = rng.uniform(500, 1500, (1, 480, 640)).astype(np.float32)
    boxes_h = [(0,) + D.get_bbox((100, 200, 220, 330)), (0,) + D.get_bbox((250, 300, 330, 420))]
    valid = torch.from_numpy(depth > 0).cuda()
    choose, ok = D.sample_choose(valid, boxes_h, 256, generator=torch.Generator(device="cuda").manual_seed(5))
    assert bool(ok.all())
    boxes = torch.tensor(boxes_h, dtype=torch.int32).cuda()
    inp = D.prepare_instances(torch.from_numpy(frames).cuda(), torch.from_numpy(depth).cuda(), boxes, choose,
                              (591.0125, 590.16775, 322.525, 244.11084), img_size=64)
    inp["category_label"] = torch.zeros(2, dtype=torch.int64, device="cuda")
    torch.manual_seed(0)
    m = M.IST_Net(6, False).cuda().eval()
    with torch.no_grad():
        ep = m(inp)
    assert ep["pred_rotation"].shape == (2, 3, 3) and torch.isfinite(ep["pred_translation"]).all()
    with pytest.raises(RuntimeError, match="CPU not supported"):
        D.prepare_instances(torch.from_numpy(frames), torch.from_numpy(depth).cuda(), boxes, choose, (1.0, 1.0, 0.0, 0.0))


def test_frames_to_poses_pipeline_equals_the_manual_steps():
    """infer.estimate_poses (device data preparation -> bucketed eval graph -> pose assembly) against the same steps done by hand."""
    from conftest import perturb_batchnorm
    from istnet_b200 import model as M
    from istnet_b200.infer import InferenceEngine, estimate_poses

    rng = np.random.default_rng(21)
    rgb = torch.from_numpy(rng.integers(0, 256, (480, 640, 3), dtype=np.uint8)).cuda()
    depth = torch.from_numpy(rng.uniform(500, 1500, (480, 640)).astype(np.float32)).cuda()
    dets = [(100, 200, 220, 330), (250, 300, 330, 420), (10, 10, 14, 13)]
    masks = torch.zeros(3, 480, 640, dtype=torch.bool, device="cuda")
    masks[0, 110:210, 210:320] = True
    masks[1, 255:325, 305:415] = True
    masks[2, 10:13, 10:13] = True  # 9 valid pixels: dropped (<= 16)
    torch.manual_seed(2)
    m = M.IST_Net(6, False)
    perturb_batchnorm(m, seed=44)
    m = m.cuda().eval()
    eng = InferenceEngine(m, npts=256, img=64)
    intr = (591.0125, 590.16775, 322.525, 244.11084)
    g = torch.Generator(device="cuda").manual_seed(9)
    rts, scales, keep = estimate_poses(eng, rgb, depth, masks, dets, [2, 5, 0], intr, generator=g)
    assert keep == [0, 1] and rts.shape == (2, 4, 4) and scales.shape == (2, 3)
    # by hand, with the same random draws
    g = torch.Generator(device="cuda").manual_seed(9)
    boxes_h = [(0,) + D.get_bbox(dets[j]) for j in keep]
    choose = torch.stack([D.sample_choose(masks[j : j + 1], [boxes_h[i]], 256, generator=g)[0][0] for i, j in enumerate(keep)])
    inp = D.prepare_instances(rgb.unsqueeze(0), depth.unsqueeze(0), torch.tensor(boxes_h, dtype=torch.int32).cuda(), choose, intr, img_size=64)
    inp["category_label"] = torch.tensor([2, 5], device="cuda")
    with torch.no_grad():
        ep = m(inp)
    s = torch.norm(ep["pred_size"], dim=1, keepdim=True)
    assert torch.allclose(rts[:, :3, :3], ep["pred_rotation"] * s.unsqueeze(2), rtol=1e-5, atol=1e-6)
    assert torch.allclose(rts[:, :3, 3], ep["pred_translation"], rtol=1e-5, atol=1e-6) and torch.allclose(scales, ep["pred_size"] / s, rtol=1e-5, atol=1e-6)
    assert torch.isfinite(rts).all()


def test_training_labels_and_augmentation_match_the_reference_goldens():
    """qo from the float64 (jittered) points in the back-projection kernel, then the bounding-box / rigid augmentations in place
    (csrc/dataprep.cu) against goldens produced by the reference's own functions (tests/golden/dataprep_aug.npz)."""
    g, a = np.load(GOLD), np.load(os.path.join(os.path.dirname(GOLD), "dataprep_aug.npz"))
    S, norm = int(g["S"]), float(g["norm_scale"])
    rot64, par = D.canonical_labels(a["rotation"], a["translation"], a["size"], a["symmetric"])
    out = D.prepare_instances(torch.from_numpy(g["frames"]).cuda(), torch.from_numpy(g["depth"]).cuda(), torch.from_numpy(g["boxes"]).cuda(),
                              torch.from_numpy(g["choose_in"]).cuda(), tuple(g["intrinsics"]), img_size=S, norm_scale=norm,
                              noise=torch.from_numpy(g["noise"]), label_params=par)
    assert np.array_equal(out["pts"].cpu().numpy(), g["pts_jitter"])
    qo = out["qo"].cpu().numpy()
    # float64 contraction of three products: BLAS and the kernel may round the last bit of the double differently; after the FP32 store
    # that is at most one ulp on a handful of elements
    assert np.allclose(qo, a["qo"], rtol=2e-7, atol=1e-9) and (qo != a["qo"]).mean() < 0.01, (np.abs(qo - a["qo"]).max(), (qo != a["qo"]).mean())
    pts_d, qo_d = out["pts"].clone(), torch.from_numpy(a["qo"]).cuda()
    Rl, tl, sl = D.augment_instances(pts_d, qo_d, a["rotation_label"], a["translation"], a["size"], a["sym0"], a["do_bb"], a["aug_bb"], a["do_rt"],
                                     a["aug_t"], a["aug_R"])
    assert np.allclose(pts_d.cpu().numpy(), a["out_pts"], rtol=2e-6, atol=2e-7), np.abs(pts_d.cpu().numpy() - a["out_pts"]).max()
    assert np.allclose(qo_d.cpu().numpy(), a["out_qo"], rtol=2e-6, atol=2e-7), np.abs(qo_d.cpu().numpy() - a["out_qo"]).max()
    assert np.allclose(Rl.numpy(), a["out_R"], rtol=2e-6, atol=2e-7) and np.allclose(tl.numpy(), a["out_t"], rtol=2e-6, atol=2e-7)
    assert np.allclose(sl.numpy(), a["out_s"], rtol=2e-6, atol=2e-7)
    # instances without augmentation are untouched bit for bit
    idle = [b for b in range(len(a["do_bb"])) if not a["do_bb"][b] and not a["do_rt"][b]]
    assert idle and all(np.array_equal(pts_d[b].cpu().numpy(), g["pts_jitter"][b]) for b in idle)
