"""Throughput of the per-instance input preparation (istnet_b200/dataprep.py): B = 32 instances (192x192 crops,
1024 points) cut out of 8 resident 480x640 frames per call, against the CPU restatement of the reference's host path on one core.
usage (GPU box): python tools/bench_dataprep.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from istnet_b200 import dataprep as D  # noqa: E402

rng = np.random.default_rng(0)
F, H, W, S, N, B = 8, 480, 640, 192, 1024, 32
frames = rng.integers(0, 256, (F, H, W, 3), dtype=np.uint8)
depth = rng.uniform(300, 2500, (F, H, W)).astype(np.float32)
boxes_h = []
for i in range(B):
    y1, x1 = int(rng.integers(0, 380)), int(rng.integers(0, 520))
    boxes_h.append((i % F,) + D.get_bbox((y1, x1, y1 + int(rng.integers(40, 100)), x1 + int(rng.integers(40, 120)))))
boxes = torch.tensor(boxes_h, dtype=torch.int32).cuda()
choose = torch.stack([torch.randint(0, (b[2] - b[1]) * (b[4] - b[3]), (N,), dtype=torch.int32) for b in boxes_h]).cuda()
fr, dp = torch.from_numpy(frames).cuda(), torch.from_numpy(depth).cuda()
intr = (591.0125, 590.16775, 322.525, 244.11084)
for _ in range(5):
    D.prepare_instances(fr, dp, boxes, choose, intr)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(200):
    D.prepare_instances(fr, dp, boxes, choose, intr)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 200
out_bytes = B * (3 * S * S * 4 + N * 12 + N * 8)
print(f"GPU: {ms * 1000:.1f} us per batch of {B} instances = {B / ms * 1000:.0f} instances/s ({out_bytes / ms / 1e6:.1f} GB/s of outputs)")
try:  # the host libraries the reference's Dataset calls, on one core (cv2.resize + torchvision transform + the numpy expressions)
    import cv2
    import torchvision.transforms as T

    tr = T.Compose([T.ToTensor(), T.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
    ch = choose.cpu().numpy().astype(np.int64)
    fx, fy, cx, cy = intr
    t0 = time.time()
    for i, b in enumerate(boxes_h):
        f, rmin, rmax, cmin, cmax = b
        tr(cv2.resize(frames[f][rmin:rmax, cmin:cmax], (S, S), interpolation=cv2.INTER_LINEAR))
        cw = cmax - cmin
        r, c = rmin + ch[i] // cw, cmin + ch[i] % cw
        z = depth[f][r, c] / 1000.0
        np.stack([(c - cx) * z / fx, (r - cy) * z / fy, z], 1).astype(np.float32)
        ratio = S / (rmax - rmin)
        (np.floor((ch[i] // (rmax - rmin)) * ratio) * S + np.floor((ch[i] % (rmax - rmin)) * ratio)).astype(np.int64)
    dt = time.time() - t0
    print(f"CPU (cv2.resize + torchvision + numpy, one core, crop / resize / normalise / back-project only): {B / dt:.0f} instances/s")
except ImportError:
    pass
