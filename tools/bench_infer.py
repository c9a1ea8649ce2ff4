"""Latency of the eval-mode inference path (istnet_b200/infer.py; reference test_func, utils/solver.py:217-241) per image for
B = 1..10 instances (1024 points + 192x192 RGB crops each), CUDA-graph replay per bucket, inputs resident on the device.
usage (GPU box): python tools/bench_infer.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from istnet_b200 import model as M  # noqa: E402
from istnet_b200.infer import InferenceEngine  # noqa: E402
from istnet_b200.synth import make_batch  # noqa: E402

torch.manual_seed(1)
m = M.IST_Net(6, False).cuda().eval()
eng = InferenceEngine(m)
for b in (1, 2, 3, 4, 6, 8, 10):
    d = make_batch(b, 1024, 192, seed=b)
    inp = {k: d[k].cuda() for k in ("rgb", "pts", "choose", "category_label")}
    for _ in range(5):
        eng(inp)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        eng(inp)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 50
    print(f"B={b:2d} (bucket {eng._bucket(b):2d}): {ms:.3f} ms per image, {b / ms * 1000:.0f} instances/s")
