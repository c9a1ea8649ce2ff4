"""NOCS mAP evaluation with the pairwise geometry on the device (reference utils/evaluation_utils.py:735-1022
`compute_independent_mAP` and the helpers it calls, :38-173 and :510-732).

The reference walks every (image, class) block and, inside it, every (prediction, ground truth) pair with Python loops: a 3-D box
overlap per pair — 20 times per pair for the rotation-symmetric classes — and a rotation / translation error per pair.  Here all
pairs of all blocks of a result list are gathered first and evaluated by TWO batched tensor computations (float64, on any device);
only the greedy assignment and the AP integration, which are sequential by definition and tiny, stay on the host.

Faithful to the reference's arithmetic, including the two things one would not write on purpose:
  * `asymmetric_3d_iou` (:121-146) reduces the transformed corner matrix [3, 8] over axis 0, so its "box" has eight extents (the
    per-corner max / min over x, y, z), and the intersection / union are products of eight numbers.  The published numbers are
    computed that way, so this module does the same;
  * overlaps are stored in float32 (:545), rotation / translation errors in float64 (:679).
Returns the same (iou_3d_aps, pose_aps) arrays; `tests/test_evaluation.py` compares them with the reference's own function on random
scenes (all six classes, symmetric ones, mugs with visible / hidden handles, misses and false positives)."""
import math

import numpy as np
import torch

SYNSET_NAMES = ["BG", "bottle", "bowl", "camera", "can", "laptop", "mug"]  # utils/evaluation_utils.py:1032-1039
_SIGNS = torch.tensor([[1, 1, 1], [1, 1, -1], [-1, 1, 1], [-1, 1, -1], [1, -1, 1], [1, -1, -1], [-1, -1, 1], [-1, -1, -1]], dtype=torch.float64)


def _corners(RT, scales):
    """get_3d_bbox + transform_coordinates_3d (:38-85) for a batch: RT [..., 4, 4], scales [..., 3] -> [..., 3, 8]."""
    box = (_SIGNS.to(scales.device) * (scales.unsqueeze(-2) / 2)).transpose(-1, -2)  # [..., 3, 8]
    hom = torch.cat([box, torch.ones_like(box[..., :1, :])], dim=-2)                  # [..., 4, 8]
    out = RT @ hom
    return out[..., :3, :] / out[..., 3:4, :]


def _iou_from_corners(b1, b2):
    """asymmetric_3d_iou (:121-146) on corner matrices [..., 3, 8] (reduction over axis 0 of the [3, 8] matrix, see module docstring)."""
    b1_max, b1_min = b1.amax(-2), b1.amin(-2)
    b2_max, b2_min = b2.amax(-2), b2.amin(-2)
    ext = torch.minimum(b1_max, b2_max) - torch.maximum(b1_min, b2_min)
    inter = torch.where(ext.amin(-1) < 0, torch.zeros_like(ext[..., 0]), ext.prod(-1))
    union = (b1_max - b1_min).prod(-1) + (b2_max - b2_min).prod(-1) - inter
    return inter / union


def pair_iou_3d(RT_1, scales_1, RT_2, scales_2, symmetric):
    """compute_3d_iou_new (:116-173) for N pairs: RT_* [N,4,4], scales_* [N,3], symmetric [N] bool (bottle / bowl / can, or a mug whose
    handle is not visible: the best of 20 rotations of box 1 about its y axis, starting from 0).  float64 in, float64 out."""
    b2 = _corners(RT_2, scales_2)
    plain = _iou_from_corners(_corners(RT_1, scales_1), b2)
    if not bool(symmetric.any()):
        return plain
    n = 20
    th = np.array([2 * math.pi * i / float(n) for i in range(n)])  # the reference's expression, evaluated the same way (:165)
    rot_h = np.zeros((n, 4, 4))
    rot_h[:, 0, 0], rot_h[:, 0, 2], rot_h[:, 2, 0], rot_h[:, 2, 2] = np.cos(th), np.sin(th), -np.sin(th), np.cos(th)
    rot_h[:, 1, 1] = rot_h[:, 3, 3] = 1.0
    rot = torch.from_numpy(rot_h).to(RT_1.device)
    sym_idx = symmetric.nonzero().squeeze(1)
    rt = RT_1[sym_idx].unsqueeze(1) @ rot.unsqueeze(0)                                   # [S, 20, 4, 4]
    ious = _iou_from_corners(_corners(rt, scales_1[sym_idx].unsqueeze(1)), b2[sym_idx].unsqueeze(1))  # [S, 20]
    # the reference folds with Python's max(max_iou, iou) from 0: a NaN candidate never replaces the running value
    best = torch.where(torch.isnan(ious), torch.zeros_like(ious), ious).amax(1).clamp_min(0.0)
    out = plain.clone()
    out[sym_idx] = best
    return out


def _det3(m):
    return (m[..., 0, 0] * (m[..., 1, 1] * m[..., 2, 2] - m[..., 1, 2] * m[..., 2, 1])
            - m[..., 0, 1] * (m[..., 1, 0] * m[..., 2, 2] - m[..., 1, 2] * m[..., 2, 0])
            + m[..., 0, 2] * (m[..., 1, 0] * m[..., 2, 1] - m[..., 1, 1] * m[..., 2, 0]))


def _cbrt(x):
    return torch.sign(x) * torch.abs(x).pow(1.0 / 3.0)


def pair_rt_errors(RT_1, RT_2, mode):
    """compute_RT_degree_cm_symmetry (:588-661) for N pairs -> [N, 2] = (rotation error in degrees, translation error in cm).
    mode [N]: 0 general, 1 symmetric about y (bottle / can / bowl, mug without visible handle), 2 y-flip classes (phone / eggbox / glue)."""
    R1 = RT_1[:, :3, :3] / _cbrt(_det3(RT_1[:, :3, :3])).view(-1, 1, 1)
    R2 = RT_2[:, :3, :3] / _cbrt(_det3(RT_2[:, :3, :3])).view(-1, 1, 1)
    y1, y2 = R1[:, :, 1], R2[:, :, 1]
    th_y = torch.acos((y1 * y2).sum(1) / (y1.norm(dim=1) * y2.norm(dim=1)))
    R = R1 @ R2.transpose(1, 2)
    tr = R.diagonal(dim1=1, dim2=2).sum(1)
    th_gen = torch.acos(((tr - 1) / 2).clamp(-1.0, 1.0))
    flip = torch.diag(torch.tensor([-1.0, 1.0, -1.0], dtype=torch.float64, device=RT_1.device))
    tr_f = (R1 @ flip @ R2.transpose(1, 2)).diagonal(dim1=1, dim2=2).sum(1)
    th_flip = torch.minimum(torch.acos((tr - 1) / 2), torch.acos((tr_f - 1) / 2))
    theta = torch.where(mode == 1, th_y, torch.where(mode == 2, th_flip, th_gen)) * (180 / math.pi)
    shift = (RT_1[:, :3, 3] - RT_2[:, :3, 3]).norm(dim=1) * 100
    return torch.stack([theta, shift], 1)


# ------------------------------------------------------------------------------------------------ host side: assignment and AP
def match_by_iou(overlaps, thresholds):
    """The assignment of compute_3d_matches (:555-585) inside one class: overlaps [P, G] with the predictions already sorted by score.
    Per threshold, every prediction takes the free ground truth of highest overlap if that overlap is ABOVE the threshold."""
    P, G = overlaps.shape
    pred_m = -np.ones((len(thresholds), P))
    gt_m = -np.ones((len(thresholds), G))
    order = [np.argsort(overlaps[i])[::-1] for i in range(P)]
    for s, thr in enumerate(thresholds):
        for i in range(P):
            cand = order[i]
            low = np.where(overlaps[i, cand] < 0)[0]
            if low.size:
                cand = cand[: low[0]]
            for j in cand:
                if gt_m[s, j] > -1:
                    continue
                if overlaps[i, j] < thr:
                    break
                if overlaps[i, j] > thr:
                    gt_m[s, j], pred_m[s, i] = i, j
                    break
    return gt_m, pred_m


def match_by_pose(errors, degree_list, shift_list):
    """compute_match_from_degree_cm (:690-732) inside one class: errors [P, G, 2]; every prediction takes the free ground truth of
    smallest (degree + cm) sum among those within both thresholds."""
    P, G = errors.shape[:2]
    pred_m = -np.ones((len(degree_list), len(shift_list), P))
    gt_m = -np.ones((len(degree_list), len(shift_list), G))
    if P == 0 or G == 0:
        return gt_m, pred_m
    order = [np.argsort(errors[i].sum(-1)) for i in range(P)]
    for d, dt in enumerate(degree_list):
        for s, st in enumerate(shift_list):
            for i in range(P):
                for j in order[i]:
                    if gt_m[d, s, j] > -1 or errors[i, j, 0] > dt or errors[i, j, 1] > st:
                        continue
                    gt_m[d, s, j], pred_m[d, s, i] = i, j
                    break
    return gt_m, pred_m


def average_precision(pred_match, pred_scores, gt_match):
    """compute_ap_from_matches_scores (:87-113): VOC-style area under the monotone precision envelope."""
    idx = np.argsort(pred_scores)[::-1]
    hit = pred_match[idx] > -1
    prec = np.cumsum(hit) / (np.arange(len(hit)) + 1)
    rec = np.cumsum(hit).astype(np.float32) / len(gt_match)
    prec = np.concatenate([[0], prec, [0]])
    rec = np.concatenate([[0], rec, [1]])
    prec = np.maximum.accumulate(prec[::-1])[::-1]
    k = np.where(rec[:-1] != rec[1:])[0] + 1
    return np.sum((rec[k] - rec[k - 1]) * prec[k])


# ------------------------------------------------------------------------------------------------ the evaluation
def compute_mAP(final_results, synset_names=SYNSET_NAMES, degree_thresholds=(360,), shift_thresholds=(100,), iou_3d_thresholds=(0.1,),
                iou_pose_thres=0.1, use_matches_for_pose=True, device=None):
    """compute_independent_mAP (:735-1022) without the plots: returns (iou_3d_aps [classes+1, n_iou], pose_aps [classes+1, n_deg+1,
    n_shift+1]) with the mean over the classes in the last row.  final_results: the dictionaries `test_func` writes
    (utils/solver.py:243-259): gt_class_ids, gt_RTs, gt_scales, gt_handle_visibility, pred_bboxes, pred_class_ids, pred_scales,
    pred_scores, pred_RTs."""
    dev = torch.device(device) if device is not None else torch.device("cuda" if torch.cuda.is_available() else "cpu")
    ncls = len(synset_names)
    deg_list, shift_list, iou_list = list(degree_thresholds) + [360], list(shift_thresholds) + [100], list(iou_3d_thresholds)
    if use_matches_for_pose:
        assert iou_pose_thres in iou_list
    sym_names, flip_names = ("bottle", "bowl", "can"), ("phone", "eggbox", "glue")

    # ---- pass 1: the (image, class) blocks and all their prediction x ground-truth pairs
    blocks, p_rt, p_sc, g_rt, g_sc, sym, mode = [], [], [], [], [], [], []
    n_pairs = 0
    for res in final_results:
        gt_ids = np.asarray(res["gt_class_ids"]).astype(np.int32)
        pr_ids = np.asarray(res["pred_class_ids"])
        if len(gt_ids) == 0 and len(pr_ids) == 0:
            continue
        gt_RTs, gt_scales, gt_hv = np.array(res["gt_RTs"]), np.array(res["gt_scales"]), np.asarray(res["gt_handle_visibility"])
        pr_RTs, pr_scales, pr_scores = np.array(res["pred_RTs"]), np.asarray(res["pred_scales"]), np.asarray(res["pred_scores"])
        for c in range(1, ncls):
            gsel = np.where(gt_ids == c)[0] if len(gt_ids) else np.zeros(0, dtype=np.int64)
            psel = np.where(pr_ids == c)[0] if len(pr_ids) else np.zeros(0, dtype=np.int64)
            scores = pr_scores[psel] if len(psel) else np.zeros(0)
            order = np.argsort(scores)[::-1] if len(psel) else np.zeros(0, dtype=np.int64)  # compute_3d_matches sorts by score (:531)
            psel = psel[order] if len(psel) else psel
            hv = gt_hv[gsel] if (synset_names[c] == "mug" and len(gsel)) else np.ones(len(gsel))
            P, G = len(psel), len(gsel)
            blocks.append({"cls": c, "P": P, "G": G, "off": n_pairs, "scores": scores[order] if P else scores, "hv": hv})
            if P and G:
                pi, gi = np.repeat(np.arange(P), G), np.tile(np.arange(G), P)
                p_rt.append(pr_RTs[psel][pi]); p_sc.append(pr_scales[psel][pi]); g_rt.append(gt_RTs[gsel][gi]); g_sc.append(gt_scales[gsel][gi])
                if synset_names[c] in sym_names:
                    sym.append(np.ones(P * G, dtype=bool))
                elif synset_names[c] == "mug":
                    sym.append(hv[gi] == 0)
                else:
                    sym.append(np.zeros(P * G, dtype=bool))
                mode.append(np.where(sym[-1], 1, 2 if synset_names[c] in flip_names else 0))
                n_pairs += P * G

    # ---- the geometry of all pairs, batched on the device
    if n_pairs:
        t = lambda a: torch.from_numpy(np.concatenate(a).astype(np.float64)).to(dev)
        RT1, S1, RT2, S2 = t(p_rt), t(p_sc), t(g_rt), t(g_sc)
        symm = torch.from_numpy(np.concatenate(sym)).to(dev)
        iou = pair_iou_3d(RT1, S1, RT2, S2, symm).cpu().numpy().astype(np.float32)           # stored in float32 like :545
        err = pair_rt_errors(RT1, RT2, torch.from_numpy(np.concatenate(mode)).to(dev)).cpu().numpy()
    else:
        iou, err = np.zeros(0, np.float32), np.zeros((0, 2))

    # ---- pass 2: assignments per block, concatenated per class
    n_iou, n_deg, n_sh = len(iou_list), len(deg_list), len(shift_list)
    iou_pm = [np.zeros((n_iou, 0)) for _ in range(ncls)]
    iou_ps = [np.zeros((n_iou, 0)) for _ in range(ncls)]
    iou_gm = [np.zeros((n_iou, 0)) for _ in range(ncls)]
    pose_pm = [np.zeros((n_deg, n_sh, 0)) for _ in range(ncls)]
    pose_ps = [np.zeros((n_deg, n_sh, 0)) for _ in range(ncls)]
    pose_gm = [np.zeros((n_deg, n_sh, 0)) for _ in range(ncls)]
    for b in blocks:
        c, P, G = b["cls"], b["P"], b["G"]
        ov = iou[b["off"] : b["off"] + P * G].reshape(P, G) if P and G else np.zeros((P, G), np.float32)
        er = err[b["off"] : b["off"] + P * G].reshape(P, G, 2) if P and G else np.zeros((P, G, 2))
        gm, pm = match_by_iou(ov, iou_list)
        iou_pm[c] = np.concatenate((iou_pm[c], pm), axis=-1)
        iou_ps[c] = np.concatenate((iou_ps[c], np.tile(b["scores"], (n_iou, 1))), axis=-1)
        iou_gm[c] = np.concatenate((iou_gm[c], gm), axis=-1)
        pk, gk = np.arange(P), np.arange(G)
        if use_matches_for_pose:  # only the pairs matched at iou_pose_thres enter the pose metric (:837-858)
            k = iou_list.index(iou_pose_thres)
            pk, gk = pk[pm[k] > -1], gk[gm[k] > -1]
        gm2, pm2 = match_by_pose(er[np.ix_(pk, gk)] if len(pk) and len(gk) else np.zeros((len(pk), len(gk), 2)), deg_list, shift_list)
        pose_pm[c] = np.concatenate((pose_pm[c], pm2), axis=-1)
        pose_ps[c] = np.concatenate((pose_ps[c], np.tile(b["scores"][pk], (n_deg, n_sh, 1))), axis=-1)
        pose_gm[c] = np.concatenate((pose_gm[c], gm2), axis=-1)

    iou_aps = np.zeros((ncls + 1, n_iou))
    pose_aps = np.zeros((ncls + 1, n_deg, n_sh))
    for c in range(1, ncls):
        for s in range(n_iou):
            iou_aps[c, s] = average_precision(iou_pm[c][s], iou_ps[c][s], iou_gm[c][s])
        for i in range(n_deg):
            for j in range(n_sh):
                pose_aps[c, i, j] = average_precision(pose_pm[c][i, j], pose_ps[c][i, j], pose_gm[c][i, j])
    iou_aps[-1] = np.mean(iou_aps[1:-1], axis=0)
    pose_aps[-1] = np.mean(pose_aps[1:-1], axis=0)
    return iou_aps, pose_aps
