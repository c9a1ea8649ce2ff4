"""Phase timeline of the graph-replayed bench step (device %globaltimer markers as graph nodes; istnet_b200/trace.py).
usage: python tools/timeline.py [--config cfg1] [--batch N]   -> prints `t_ms  stream  phase` sorted by time."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from istnet_b200 import trace  # noqa: E402
from istnet_b200.graph import GraphedTrainStep  # noqa: E402
from istnet_b200.synth import make_batch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="cfg1")
ap.add_argument("--batch", type=int, default=0)
a = ap.parse_args()
wl = dict(bench.WORKLOADS[a.config])
if a.batch:
    wl["batch"] = a.batch
dev = torch.device("cuda", 0)
model, loss_fn = bench.build_model(wl["model"], dev)
data = {k: v.to(dev) for k, v in make_batch(wl["batch"], wl["npts"], wl["img"], seed=1).items()}


class _Step(GraphedTrainStep):
    def _eager(self):
        trace.reset()
        trace.mark("step start")
        out = super()._eager()
        trace.mark("step end")
        return out


trace.enable()
step = _Step(model, loss_fn, data, bench.MODEL_IN, bench.LABELS)
for _ in range(5):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
step()
e1.record()
torch.cuda.synchronize()
print(f"graph replay with markers: {e0.elapsed_time(e1):.2f} ms")
for t, sid, name in trace.report():
    print(f"{t:8.3f} ms  s{sid}  {name}")
