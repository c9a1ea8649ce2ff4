"""Per-instance input preparation on the device.

The reference prepares every instance on the host, inside `Dataset.__getitem__` (provider/dataset.py:162-233 for training,
:333-433 for testing): crop + `cv2.resize` + `ToTensor` / `Normalize`, back-projection of the chosen depth pixels, `choose`
re-mapping — per instance, in `num_workers: 4` processes; at > 1200 instances/s per GPU that is the end-to-end limit.  Here the
decoded frames (RGB uint8 + hole-filled depth float32) are uploaded once and ONE call prepares all instances of a batch with the
kernels of csrc/dataprep.cu, bit-exact with OpenCV / torchvision / numpy (tests/test_gpu_dataprep.py).

Out of this module's scope (stay on the host, as in the reference): image decoding, `fill_missing` (OpenCV morphology + bilateral
filter, utils/data_utils.py:357-534), ColorJitter (PIL), the shape / pose augmentations of the labels."""
import ctypes

import torch

MEAN = (0.485, 0.456, 0.406)  # provider/dataset.py:69-72
STD = (0.229, 0.224, 0.225)


def get_bbox(bbox, img_h=480, img_w=640):
    """utils/data_utils.py:43-71: square crop window (multiple of 40, <= 440) around a detection box (y1, x1, y2, x2)."""
    y1, x1, y2, x2 = (int(v) for v in bbox)
    window = min((max(y2 - y1, x2 - x1) // 40 + 1) * 40, 440)
    cy, cx = (y1 + y2) // 2, (x1 + x2) // 2
    rmin, rmax, cmin, cmax = cy - window // 2, cy + window // 2, cx - window // 2, cx + window // 2
    if rmin < 0:
        rmin, rmax = 0, rmax - rmin
    if cmin < 0:
        cmin, cmax = 0, cmax - cmin
    if rmax > img_h:
        rmin, rmax = rmin - (rmax - img_h), img_h
    if cmax > img_w:
        cmin, cmax = cmin - (cmax - img_w), img_w
    return rmin, rmax, cmin, cmax


def sample_choose(valid_mask, boxes, n, generator=None):
    """dataset.py:192-200 on the device: for every instance, `n` indices into its flattened crop drawn from the valid pixels
    (mask & depth > 0) — without replacement when there are more than `n`, with replacement otherwise.  valid_mask [F,H,W] bool,
    boxes [B,5] (frame, rmin, rmax, cmin, cmax) on the host.  Returns int32 [B,n] and a bool [B] (False: no valid pixel, the
    reference then re-draws another sample).  torch's generator, not numpy's: statistically the same draw, not the same numbers."""
    out = torch.zeros(len(boxes), n, dtype=torch.int32, device=valid_mask.device)
    ok = torch.zeros(len(boxes), dtype=torch.bool)
    for i, (f, rmin, rmax, cmin, cmax) in enumerate(boxes):
        idx = valid_mask[f, rmin:rmax, cmin:cmax].reshape(-1).nonzero().squeeze(1)
        if idx.numel() == 0:
            continue
        if idx.numel() <= n:
            pick = torch.randint(idx.numel(), (n,), device=idx.device, generator=generator)
        else:
            pick = torch.randperm(idx.numel(), device=idx.device, generator=generator)[:n]
        out[i] = idx[pick].to(torch.int32)
        ok[i] = True
    return out, ok


def canonical_labels(rotation, translation, size, symmetric):
    """The per-instance part of dataset.py:236-250 on the host (3x3 work): for the rotation-symmetric classes (`cat_id in sym_ids`)
    the rotation label is re-based so that its x axis carries no rotation about y (float64 from there on, as in the reference).
    Returns (rotation [B,3,3] float64, label parameter block [B,13] float64 = t, |size| + 1e-8, R) for prepare_instances(labels=...)."""
    import math

    import numpy as np

    rotation = np.asarray(rotation, dtype=np.float32)
    translation, size = np.asarray(translation, dtype=np.float32), np.asarray(size, dtype=np.float32)
    B = rotation.shape[0]
    rot64, par = np.zeros((B, 3, 3)), np.zeros((B, 13))
    for b in range(B):
        r = rotation[b]
        if bool(symmetric[b]):
            tx, ty = r[0, 0] + r[2, 2], r[0, 2] - r[2, 0]
            rn = math.sqrt(tx ** 2 + ty ** 2)
            r = r @ np.array([[tx / rn, 0.0, -ty / rn], [0.0, 1.0, 0.0], [ty / rn, 0.0, tx / rn]])
        rot64[b] = r
        par[b, 0:3], par[b, 3], par[b, 4:13] = translation[b], np.linalg.norm(size[b]) + 1e-8, np.asarray(r, dtype=np.float64).reshape(9)
    return rot64, par


def augment_instances(pts, qo, rotation, translation, size, sym0, do_bb, aug_bb, do_rt, aug_t, aug_R):
    """provider/data_augmentation.py:208-233 with the default configuration (bounding-box deformation and rigid perturbation; the
    Bernoulli draws do_bb / do_rt [B] and the parameters of dataset.py:124-136 come from the caller's RNG): the points and NOCS
    coordinates [B,N,3] (CUDA, modified in place) by one kernel, the 3x3 / 3-vector labels on the host in FP32 like the torch CPU
    code.  rotation [B,3,3], translation / size [B,3], sym0 [B] (sym_info[0]), aug_bb / aug_t [B,3], aug_R [B,3,3]: host arrays.
    Returns (rotation, translation, size) float32 tensors."""
    from . import _C
    from ._C import c_int, ptr

    if not (pts.is_cuda and (qo is None or qo.is_cuda)):
        raise RuntimeError("CPU not supported")
    B, N, _ = pts.shape
    par, R, t, s = augment_label_params(rotation, translation, size, sym0, do_bb, aug_bb, do_rt, aug_t, aug_R)
    par_d = par.to(pts.device)
    _C.call("augment_points", c_int(B), c_int(N), ptr(par_d), ptr(pts), ptr(qo) if qo is not None else None)
    return R, t, s


def augment_label_params(rotation, translation, size, sym0, do_bb, aug_bb, do_rt, aug_t, aug_R):
    """Host half of augment_instances: the kernel's parameter block [B,32] and the augmented (rotation, translation, size) labels."""
    B = len(do_bb)
    R = torch.as_tensor(rotation, dtype=torch.float32).clone().reshape(B, 3, 3)
    t = torch.as_tensor(translation, dtype=torch.float32).clone().reshape(B, 3)
    s = torch.as_tensor(size, dtype=torch.float32).clone().reshape(B, 3)
    par = torch.zeros(B, 32, dtype=torch.float32)
    for b in range(B):
        par[b, 0:9], par[b, 9:12] = R[b].reshape(9), t[b]
        e = torch.ones(3)
        k = torch.tensor(1.0)
        if bool(do_bb[b]):
            e = torch.as_tensor(aug_bb[b], dtype=torch.float32).clone()
            if int(sym0[b]) == 1:  # y-axis symmetry: x and z stretch together (:49-56)
                e[0] = e[2] = (e[0] + e[2]) / 2
            k = torch.norm(torch.tensor([s[b, 0] * e[0], s[b, 1] * e[1], s[b, 2] * e[2]])) / torch.norm(s[b])
            s[b] = s[b] * e
        par[b, 12:15], par[b, 15], par[b, 16] = e, k, 1.0 if bool(do_bb[b]) else 0.0
        if bool(do_rt[b]):
            d = torch.as_tensor(aug_t[b], dtype=torch.float32)
            Rm = torch.as_tensor(aug_R[b], dtype=torch.float32)
            par[b, 17:20], par[b, 20:29], par[b, 29] = d, Rm.reshape(9), 1.0
            t[b] = torch.mm(Rm, (t[b] + d).view(3, 1)).view(3)
            R[b] = torch.mm(Rm, R[b])
    return par, R, t, s


def prepare_instances(rgb_frames, depth, boxes, choose, intrinsics, img_size=192, norm_scale=1000.0, noise=None, label_params=None):
    """rgb_frames [F,H,W,3] uint8 (RGB), depth [F,H,W] float32, boxes [B,5] int32 (frame, rmin, rmax, cmin, cmax), choose [B,N]
    int32 crop-pixel indices — all CUDA tensors; intrinsics (fx, fy, cx, cy); noise: optional float64 [B,N,3] jitter
    (dataset.py:210); label_params: optional [B,13] float64 from canonical_labels (training: adds 'qo' [B,N,3], dataset.py:249).
    Returns the model's inputs {'rgb' [B,3,S,S] f32, 'pts' [B,N,3] f32, 'choose' [B,N] int64}."""
    from . import _C
    from ._C import c_float, c_int, ptr

    if not (rgb_frames.is_cuda and depth.is_cuda and boxes.is_cuda and choose.is_cuda):
        raise RuntimeError("CPU not supported")
    if rgb_frames.dtype != torch.uint8 or depth.dtype != torch.float32 or boxes.dtype != torch.int32 or choose.dtype != torch.int32:
        raise TypeError("prepare_instances: rgb_frames uint8, depth float32, boxes / choose int32")
    rgb_frames, depth, boxes, choose = rgb_frames.contiguous(), depth.contiguous(), boxes.contiguous(), choose.contiguous()
    F, H, W, _ = rgb_frames.shape
    B, N = choose.shape
    dev = rgb_frames.device
    rgb = torch.empty(B, 3, img_size, img_size, dtype=torch.float32, device=dev)
    pts = torch.empty(B, N, 3, dtype=torch.float32, device=dev)
    cho = torch.empty(B, N, dtype=torch.int64, device=dev)
    if noise is not None:
        noise = noise.to(device=dev, dtype=torch.float64).contiguous()
    qo = lab = None
    if label_params is not None:
        lab = torch.as_tensor(label_params, dtype=torch.float64).to(dev).contiguous()
        qo = torch.empty(B, N, 3, dtype=torch.float32, device=dev)
    fx, fy, cx, cy = (float(v) for v in intrinsics)
    mean, std = (ctypes.c_float * 3)(*MEAN), (ctypes.c_float * 3)(*STD)
    _C.call("prepare_instances", ptr(rgb_frames), ptr(depth), c_int(F), c_int(H), c_int(W), ptr(boxes), ptr(choose), c_int(B), c_int(N),
            c_int(img_size), ctypes.c_double(fx), ctypes.c_double(fy), ctypes.c_double(cx), ctypes.c_double(cy), c_float(norm_scale), mean, std,
            ptr(noise) if noise is not None else None, ptr(lab) if lab is not None else None, ptr(rgb), ptr(pts),
            ptr(qo) if qo is not None else None, ptr(cho))
    out = {"rgb": rgb, "pts": pts, "choose": cho}
    if qo is not None:
        out["qo"] = qo
    return out
