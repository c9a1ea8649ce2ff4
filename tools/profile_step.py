"""Runs W warm-up steps and then K profiled training steps of the bench workload between cudaProfilerStart/Stop
(use with `ncu --profile-from-start off`).  Numbers printed under a profiler are never bench values."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from istnet_b200.synth import make_batch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="cfg1")
ap.add_argument("--batch", type=int, default=0)
ap.add_argument("--warmup", type=int, default=2)
ap.add_argument("--steps", type=int, default=1)
a = ap.parse_args()
wl = dict(bench.WORKLOADS[a.config])
if a.batch:
    wl["batch"] = a.batch
dev = torch.device("cuda", 0)
model, loss_fn = bench.build_model(wl["model"], dev)
data = {k: v.to(dev) for k, v in make_batch(wl["batch"], wl["npts"], wl["img"], seed=1).items()}


def step():
    model.zero_grad(set_to_none=True)
    ep = model({k: data[k] for k in bench.MODEL_IN})
    ep.update({k: data[k] for k in bench.LABELS})
    loss_fn(ep).backward()


for _ in range(a.warmup):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(a.steps):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
