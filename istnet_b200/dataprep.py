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


def prepare_instances(rgb_frames, depth, boxes, choose, intrinsics, img_size=192, norm_scale=1000.0, noise=None):
    """rgb_frames [F,H,W,3] uint8 (RGB), depth [F,H,W] float32, boxes [B,5] int32 (frame, rmin, rmax, cmin, cmax), choose [B,N]
    int32 crop-pixel indices — all CUDA tensors; intrinsics (fx, fy, cx, cy); noise: optional float64 [B,N,3] jitter
    (dataset.py:210).  Returns the model's inputs {'rgb' [B,3,S,S] f32, 'pts' [B,N,3] f32, 'choose' [B,N] int64}."""
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
    fx, fy, cx, cy = (float(v) for v in intrinsics)
    mean, std = (ctypes.c_float * 3)(*MEAN), (ctypes.c_float * 3)(*STD)
    _C.call("prepare_instances", ptr(rgb_frames), ptr(depth), c_int(F), c_int(H), c_int(W), ptr(boxes), ptr(choose), c_int(B), c_int(N),
            c_int(img_size), ctypes.c_double(fx), ctypes.c_double(fy), ctypes.c_double(cx), ctypes.c_double(cy), c_float(norm_scale), mean, std,
            ptr(noise) if noise is not None else None, ptr(rgb), ptr(pts), ptr(cho))
    return {"rgb": rgb, "pts": pts, "choose": cho}
