"""istnet_b200.evaluation (pairwise 3-D box overlap / pose errors batched on a device, greedy assignment and AP
on the host) against the reference's own `compute_independent_mAP` and helpers (utils/evaluation_utils.py) on random scenes.
Runs where /root/reference exists; a committed golden (tests/golden/evaluation.npz, produced by the reference here) covers the GPU box."""
import os
import sys

import numpy as np
import pytest
import torch

from istnet_b200 import evaluation as E

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
GOLD = os.path.join(ROOT, "tests", "golden", "evaluation.npz")
DEG, SHIFT, IOU = [5, 10, 15], [2, 5, 10], [0.1, 0.25, 0.5, 0.75]


def _rot(rng):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def _small_rot(rng, deg):
    ax = rng.normal(size=3)
    ax /= np.linalg.norm(ax)
    a = np.deg2rad(deg)
    K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
    return np.eye(3) + np.sin(a) * K + (1 - np.cos(a)) * K @ K


def make_results(seed, n_images=25):
    """Result dictionaries in the layout test_func writes (utils/solver.py:243-259): similarity transforms (scaled rotations),
    predictions = perturbed ground truths + misses + false positives, all six classes, mugs with and without visible handles."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n_images):
        g = int(rng.integers(0, 6))
        ids = rng.integers(1, 7, g)
        RTs, scales, hv = np.zeros((g, 4, 4)), np.zeros((g, 3)), np.ones(g, dtype=np.int32)
        for j in range(g):
            s = rng.uniform(0.1, 0.4)
            RTs[j, :3, :3] = s * _rot(rng)
            RTs[j, :3, 3] = rng.uniform(-0.3, 0.3, 3) + np.array([0, 0, 1.0])
            RTs[j, 3, 3] = 1
            scales[j] = rng.uniform(0.3, 1.0, 3)
            hv[j] = int(rng.integers(0, 2)) if ids[j] == 6 else 1
        p_ids, p_RTs, p_scales, p_scores = [], [], [], []
        for j in range(g):
            if rng.uniform() < 0.15:
                continue  # missed
            RT = RTs[j].copy()
            RT[:3, :3] = _small_rot(rng, rng.choice([1.0, 4.0, 8.0, 20.0, 60.0])) @ RT[:3, :3] * rng.uniform(0.9, 1.1)
            RT[:3, 3] += rng.normal(0, rng.choice([0.005, 0.02, 0.06]), 3)
            p_ids.append(ids[j] if rng.uniform() > 0.1 else int(rng.integers(1, 7)))
            p_RTs.append(RT); p_scales.append(scales[j] * rng.uniform(0.85, 1.15, 3)); p_scores.append(rng.uniform(0.3, 1.0))
        for _ in range(int(rng.integers(0, 3))):  # false positives
            RT = np.eye(4); RT[:3, :3] = rng.uniform(0.1, 0.4) * _rot(rng); RT[:3, 3] = rng.uniform(-0.3, 0.3, 3) + np.array([0, 0, 1.0])
            p_ids.append(int(rng.integers(1, 7))); p_RTs.append(RT); p_scales.append(rng.uniform(0.3, 1.0, 3)); p_scores.append(rng.uniform(0.1, 0.9))
        n = len(p_ids)
        out.append({"gt_class_ids": ids.astype(np.int32), "gt_RTs": RTs, "gt_scales": scales, "gt_handle_visibility": hv,
                    "gt_bboxes": rng.integers(1, 400, (g, 4)), "pred_class_ids": np.array(p_ids, dtype=np.int32),
                    "pred_RTs": np.array(p_RTs).reshape(n, 4, 4), "pred_scales": np.array(p_scales).reshape(n, 3),
                    "pred_scores": np.array(p_scores), "pred_bboxes": rng.integers(1, 400, (n, 4))})
    return out


@pytest.fixture()
def ref_eval(tmp_path):
    if not os.path.isdir(os.path.join(REF, "utils")):
        pytest.skip("reference tree not present")
    saved = list(sys.path)
    sys.path.insert(0, os.path.join(ROOT, "compat"))
    sys.path.append(os.path.join(REF, "utils"))
    import evaluation_utils as ref
    yield ref, str(tmp_path)
    sys.path[:] = saved
    for k in [k for k in sys.modules if k.split(".")[0] in ("evaluation_utils", "matplotlib")]:
        sys.modules.pop(k, None)


def test_pair_geometry_matches_the_reference_functions(ref_eval):
    ref, _ = ref_eval
    rng = np.random.default_rng(5)
    res = make_results(11, 12)
    names = E.SYNSET_NAMES
    for r in res:
        for i in range(len(r["pred_class_ids"])):
            for j in range(len(r["gt_class_ids"])):
                c = int(r["gt_class_ids"][j])
                hv = int(r["gt_handle_visibility"][j])
                want_iou = ref.compute_3d_iou_new(r["pred_RTs"][i], r["gt_RTs"][j], r["pred_scales"][i], r["gt_scales"][j], hv, names[c], names[c])
                want_err = ref.compute_RT_degree_cm_symmetry(r["pred_RTs"][i], r["gt_RTs"][j], c, hv, names)
                symm = names[c] in ("bottle", "bowl", "can") or (names[c] == "mug" and hv == 0)
                t = lambda a: torch.from_numpy(np.asarray(a, dtype=np.float64)).unsqueeze(0)
                got_iou = E.pair_iou_3d(t(r["pred_RTs"][i]), t(r["pred_scales"][i]), t(r["gt_RTs"][j]), t(r["gt_scales"][j]), torch.tensor([symm]))
                got_err = E.pair_rt_errors(t(r["pred_RTs"][i]), t(r["gt_RTs"][j]), torch.tensor([1 if symm else 0]))
                assert abs(float(got_iou) - want_iou) <= 1e-12 * max(1.0, abs(want_iou)), (c, hv, float(got_iou), want_iou)
                assert np.allclose(got_err.numpy()[0], want_err, rtol=1e-9, atol=1e-9), (c, hv, got_err, want_err)
    assert rng is not None


def test_mAP_equals_the_reference_on_random_scenes(ref_eval):
    ref, tmp = ref_eval
    for seed in (1, 2):
        res = make_results(seed)
        want_iou, want_pose = ref.compute_independent_mAP(res, E.SYNSET_NAMES, degree_thresholds=DEG, shift_thresholds=SHIFT,
                                                          iou_3d_thresholds=IOU, iou_pose_thres=0.1, use_matches_for_pose=True,
                                                          plot_figure=False, log_dir=tmp)
        got_iou, got_pose = E.compute_mAP(res, E.SYNSET_NAMES, DEG, SHIFT, IOU, 0.1, True, device="cpu")
        assert np.array_equal(got_iou, want_iou), np.abs(got_iou - want_iou).max()
        assert np.array_equal(got_pose, want_pose), np.abs(got_pose - want_pose).max()
        assert 0.05 < got_iou[-1, 1] < 1.0 and 0.0 < got_pose[-1, 1, 1] < 1.0  # a non-degenerate test: some hits, some misses
    if not os.path.exists(GOLD) or os.environ.get("ISTNET_WRITE_GOLDEN") == "1":
        res = make_results(3)
        wi, wp = ref.compute_independent_mAP(res, E.SYNSET_NAMES, degree_thresholds=DEG, shift_thresholds=SHIFT, iou_3d_thresholds=IOU,
                                             iou_pose_thres=0.1, use_matches_for_pose=True, plot_figure=False, log_dir=tmp)
        np.savez_compressed(GOLD, iou_aps=wi, pose_aps=wp)


def test_mAP_matches_the_committed_reference_golden():
    if not os.path.exists(GOLD):
        pytest.skip("golden not generated yet")
    g = np.load(GOLD)
    got_iou, got_pose = E.compute_mAP(make_results(3), E.SYNSET_NAMES, DEG, SHIFT, IOU, 0.1, True, device="cpu")
    assert np.array_equal(got_iou, g["iou_aps"]) and np.array_equal(got_pose, g["pose_aps"])


@pytest.mark.gpu
def test_mAP_on_the_gpu_matches_the_committed_reference_golden():
    g = np.load(GOLD)
    got_iou, got_pose = E.compute_mAP(make_results(3), E.SYNSET_NAMES, DEG, SHIFT, IOU, 0.1, True, device="cuda")
    # float64 on the device: FMA contraction may move an overlap by one ulp before the float32 store; the APs are step functions of it
    assert np.abs(got_iou - g["iou_aps"]).max() <= 1e-9 and np.abs(got_pose - g["pose_aps"]).max() <= 1e-9
