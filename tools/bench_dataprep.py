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
try:
    import cv2

    from oracle import dataprep_ref as R  # noqa: E402  (CPU comparison leg only)

    ch = choose.cpu().numpy().astype(np.int64)
    t0 = time.time()
    for i, b in enumerate(boxes_h):
        crop = frames[b[0]][b[1]:b[2], b[3]:b[4]]
        R.normalize_u8(cv2.resize(crop, (S, S), interpolation=cv2.INTER_LINEAR))
        R.back_project(depth[b[0]], ch[i], b[1:], intr, 1000.0)
        R.remap_choose(ch[i], b[1:], S)
    dt = time.time() - t0
    print(f"CPU (cv2.resize + numpy, one core, crop / resize / normalise / back-project only): {B / dt:.0f} instances/s")
except ImportError:
    pass
