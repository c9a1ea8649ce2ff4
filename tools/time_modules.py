"""Per-module forward+backward time (CUDA-graph replay, B=32 bench shapes): where does the step go?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from istnet_b200 import model as M
from istnet_b200.image import ModifiedResnet
from istnet_b200.pointnet2 import PointNet2MSG
from istnet_b200.synth import make_batch

dev = "cuda"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
d = {k: v.to(dev) for k, v in make_batch(B, 1024, 192, seed=1).items()}
pts = (d["pts"] - d["pts"].mean(1, keepdim=True)).contiguous()

def graph_time(fn, params, n=10):
    for _ in range(3):
        for p in params: p.grad = None
        fn()
    for p in params:
        if p.grad is not None: p.grad = torch.zeros_like(p.grad)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g):
        fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

torch.manual_seed(1)
img = ModifiedResnet().to(dev).train()
def f_img():
    out = img.gather_rows(d["rgb"], d["choose"]); out.square().mean().backward()
print(f"image branch fwd+bwd      : {graph_time(f_img, list(img.parameters())):.2f} ms")
def f_img_fwd():
    with torch.no_grad(): img.gather_rows(d["rgb"], d["choose"])
print(f"image branch fwd only     : {graph_time(f_img_fwd, []):.2f} ms")
pn = PointNet2MSG(M.CAM_RADII).to(dev).train()
def f_pn():
    out = pn.forward_rows(pts); out.square().mean().backward()
print(f"PointNet2MSG fwd+bwd (x1) : {graph_time(f_pn, list(pn.parameters())):.2f} ms")
def f_pn_fwd():
    with torch.no_grad(): pn.forward_rows(pts)
print(f"PointNet2MSG fwd only     : {graph_time(f_pn_fwd, []):.2f} ms")
he = M.HeavyEstimator().to(dev).train()
f128 = torch.randn(B, 1024, 128, device=dev, requires_grad=True); f128b = torch.randn(B, 1024, 128, device=dev, requires_grad=True); f128c = torch.randn(B, 1024, 128, device=dev, requires_grad=True)
def f_he():
    r, t, s = he(pts, d["qo"], f128, f128b, f128c); (r.sum() + t.sum() + s.sum()).backward()
print(f"HeavyEstimator fwd+bwd    : {graph_time(f_he, list(he.parameters())):.2f} ms")
fd = M.FeatureDeformer().to(dev).train()
cls = d["category_label"].reshape(-1)
def f_fd():
    x, q = fd(pts, f128, f128b, cls); (x.sum() + q.sum()).backward()
print(f"FeatureDeformer fwd+bwd   : {graph_time(f_fd, list(fd.parameters())):.2f} ms")
