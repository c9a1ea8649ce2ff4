"""Golden vectors for the per-instance input preparation, produced by the libraries the reference itself calls —
cv2.resize / torchvision ToTensor + Normalize / numpy — with the expressions of provider/dataset.py:186-233 (train path; the test
path :369-409 is the same arithmetic without the jitter).  Run here (cv2 + torchvision present):  python tests/gen_golden_dataprep.py
-> tests/golden/dataprep.npz.  The frames are small synthetic ones (the arithmetic does not depend on the frame size)."""
import os
import sys

import cv2
import numpy as np
import torch
import torchvision.transforms as transforms

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from istnet_b200.dataprep import get_bbox  # noqa: E402

H, W, S, N, F = 120, 160, 48, 64, 2
INTR = [591.0125, 590.16775, 322.525 / 4, 244.11084 / 4]  # the Real intrinsics (dataset.py:39) with the principal point scaled to the frame
NORM = 1000.0


def main():
    rng = np.random.default_rng(7)
    yy, xx = np.mgrid[0:H, 0:W]
    frames = np.stack([np.clip(np.stack([127 + 90 * np.sin(xx / (5.0 + f) + c) * np.cos(yy / (7.0 + c)) for c in range(3)], -1)
                               + rng.normal(0, 12, (H, W, 3)), 0, 255).astype(np.uint8) for f in range(F)])
    depth = (600 + 300 * np.sin(xx / 23.0)[None] * np.cos(yy / 17.0)[None] + rng.normal(0, 3, (F, H, W))).astype(np.float32)
    depth[:, 10:14, 20:40] = 0
    # windows: squares of several sizes (up- and down-sampling to S), touching the borders, one per instance
    boxes = []
    for (f, y1, x1, y2, x2) in ((0, 10, 20, 50, 70), (0, 0, 0, 30, 25), (1, 60, 90, 118, 158), (1, 40, 40, 47, 49), (0, 5, 100, 110, 150), (1, 30, 10, 95, 60)):
        win = min((max(y2 - y1, x2 - x1) // 8 + 1) * 8, 104)   # get_bbox's rule at 1/5 scale (multiples of 8, <= 104) so that the windows fit
        cy_, cx_ = (y1 + y2) // 2, (x1 + x2) // 2
        rmin, rmax, cmin, cmax = cy_ - win // 2, cy_ + win // 2, cx_ - win // 2, cx_ + win // 2
        if rmin < 0: rmin, rmax = 0, rmax - rmin
        if cmin < 0: cmin, cmax = 0, cmax - cmin
        if rmax > H: rmin, rmax = rmin - (rmax - H), H
        if cmax > W: cmin, cmax = cmin - (cmax - W), W
        boxes.append((f, rmin, rmax, cmin, cmax))
    boxes = np.array(boxes, dtype=np.int32)
    transform = transforms.Compose([transforms.ToTensor(), transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
    xmap = np.array([[i for i in range(W)] for j in range(H)])
    ymap = np.array([[j for i in range(W)] for j in range(H)])
    cam_fx, cam_fy, cam_cx, cam_cy = INTR
    rgbs, ptss, ptsn, chos, chin, noises = [], [], [], [], [], []
    for (f, rmin, rmax, cmin, cmax) in boxes:
        d = depth[f]
        mask = d > 0
        choose = mask[rmin:rmax, cmin:cmax].flatten().nonzero()[0]
        choose = choose[rng.choice(len(choose), N, replace=len(choose) <= N)]
        pts2 = d.copy() / NORM
        pts0 = (xmap - cam_cx) * pts2 / cam_fx
        pts1 = (ymap - cam_cy) * pts2 / cam_fy
        pts = np.transpose(np.stack([pts0, pts1, pts2]), (1, 2, 0)).astype(np.float32)
        pts = pts[rmin:rmax, cmin:cmax, :].reshape((-1, 3))[choose, :]
        noise = np.clip(0.001 * rng.standard_normal((pts.shape[0], 3)), -0.005, 0.005)
        rgb = frames[f][rmin:rmax, cmin:cmax, :]
        rgb = cv2.resize(rgb, (S, S), interpolation=cv2.INTER_LINEAR)
        rgb = transform(np.array(rgb))
        crop_w = rmax - rmin
        ratio = S / crop_w
        cho = (np.floor((choose // crop_w) * ratio) * S + np.floor((choose % crop_w) * ratio)).astype(np.int64)
        rgbs.append(torch.FloatTensor(rgb).numpy()); ptss.append(torch.FloatTensor(pts).numpy()); ptsn.append(torch.FloatTensor(pts + noise).numpy())
        chos.append(cho); chin.append(choose.astype(np.int32)); noises.append(noise)
    out = os.path.join(HERE, "golden", "dataprep.npz")
    np.savez_compressed(out, frames=frames, depth=depth, boxes=boxes, choose_in=np.stack(chin), noise=np.stack(noises), intrinsics=np.array(INTR),
                        rgb=np.stack(rgbs), pts=np.stack(ptss), pts_jitter=np.stack(ptsn), choose=np.stack(chos), S=S, norm_scale=NORM)
    print(out, os.path.getsize(out), "bytes; windows:", [tuple(int(v) for v in b) for b in boxes])
    # ---- training labels (dataset.py:236-250) and the two default augmentations (provider/data_augmentation.py:45-130), produced by the
    # reference's own functions
    import importlib.util as ilu
    import math
    aug_path = "/root/reference/provider/data_augmentation.py"
    if os.path.isfile(aug_path):
        spec = ilu.spec_from_file_location("ref_aug", aug_path)
        A = ilu.module_from_spec(spec); spec.loader.exec_module(A)
        B = len(boxes)
        rot_in, tr_in, size_in = np.zeros((B, 3, 3), np.float32), np.zeros((B, 3), np.float32), np.zeros((B, 3), np.float32)
        symmetric = np.array([1, 0, 1, 0, 0, 1], dtype=bool)
        sym0 = np.array([1, 0, 1, 0, 0, 0])            # sym_info[0] of get_sym_info (mug with handle: 0)
        do_bb, do_rt = np.array([1, 1, 0, 0, 1, 1], bool), np.array([1, 0, 1, 0, 1, 1], bool)
        aug_bb = rng.uniform(0.8, 1.2, (B, 3)).astype(np.float32)
        aug_t = (rng.uniform(-50, 50, (B, 3)) / 1000.0).astype(np.float32)
        aug_R = np.stack([A.get_rotation(*rng.uniform(-15, 15, 3)) for _ in range(B)])
        qo_l, rot_l, out = [], [], {k: [] for k in ("pts", "qo", "R", "t", "s")}
        for b in range(B):
            rotation = np.linalg.qr(rng.normal(size=(3, 3)))[0].astype(np.float32)
            translation = np.array([0.02, -0.03, 0.75], np.float32) + rng.normal(0, 0.02, 3).astype(np.float32)
            size = rng.uniform(0.1, 0.35, 3).astype(np.float32)
            rot_in[b], tr_in[b], size_in[b] = rotation, translation, size
            pts = ptss[b].astype(np.float32) + noises[b]                     # float32 points + float64 jitter (dataset.py:209-210)
            if symmetric[b]:                                                # dataset.py:241-248
                theta_x = rotation[0, 0] + rotation[2, 2]
                theta_y = rotation[0, 2] - rotation[2, 0]
                r_norm = math.sqrt(theta_x**2 + theta_y**2)
                s_map = np.array([[theta_x/r_norm, 0.0, -theta_y/r_norm], [0.0, 1.0, 0.0], [theta_y/r_norm, 0.0, theta_x/r_norm]])
                rotation = rotation @ s_map
            qo = (pts - translation[np.newaxis, :]) / (np.linalg.norm(size)+1e-8) @ rotation
            qo_l.append(torch.FloatTensor(qo).numpy()); rot_l.append(torch.FloatTensor(rotation).numpy())
            PC, RR, TT, SS, NN = torch.FloatTensor(pts), torch.FloatTensor(rotation), torch.FloatTensor(translation), torch.FloatTensor(size), torch.FloatTensor(qo)
            sym_info = np.array([sym0[b], 1, 0, 1])
            if do_bb[b]:
                PC, SS, NN, _ = A.defor_3D_bb(PC, RR, TT, SS, NN, torch.zeros(8, 3), sym=sym_info, aug_bb=torch.as_tensor(aug_bb[b]))
            if do_rt[b]:
                PC, RR, TT = A.defor_3D_rt(PC, RR, TT, torch.as_tensor(aug_t[b]), torch.as_tensor(aug_R[b]))
                TT = TT.view(-1)
            for k, v in zip(("pts", "qo", "R", "t", "s"), (PC, NN, RR, TT, SS)):
                out[k].append(v.numpy().copy())
        out2 = os.path.join(HERE, "golden", "dataprep_aug.npz")
        np.savez_compressed(out2, rotation=rot_in, translation=tr_in, size=size_in, symmetric=symmetric, sym0=sym0, do_bb=do_bb, do_rt=do_rt,
                            aug_bb=aug_bb, aug_t=aug_t, aug_R=aug_R, qo=np.stack(qo_l), rotation_label=np.stack(rot_l),
                            **{"out_" + k: np.stack(v) for k, v in out.items()})
        print(out2, os.path.getsize(out2), "bytes")
    # get_bbox restatement against the reference's own function on random detection boxes
    ref_utils = "/root/reference/utils"
    if os.path.isdir(ref_utils):
        import importlib.util
        spec = importlib.util.spec_from_file_location("ref_data_utils", os.path.join(ref_utils, "data_utils.py"))
        mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
        for _ in range(2000):
            y1, x1 = int(rng.integers(0, 470)), int(rng.integers(0, 630))
            b = (y1, x1, int(rng.integers(y1 + 1, 481)), int(rng.integers(x1 + 1, 641)))
            assert tuple(mod.get_bbox(b)) == get_bbox(b), b
        print("get_bbox == reference on 2000 random boxes")


main()
