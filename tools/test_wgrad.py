"""GPU scratch test of the tcgen05 wgrad kernel against an FP64 torch reference (run under `timeout`)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from istnet_b200 import tc

torch.manual_seed(0)
dev = "cuda"

def check(B, H, W, cin, cout, k):
    x = torch.randn(B, cin, H, W, device=dev, dtype=torch.float64)
    w = torch.randn(cout, cin, k, k, device=dev, dtype=torch.float64, requires_grad=True)
    y = F.conv2d(x, w, None, padding=k // 2)
    dy = torch.randn_like(y)
    (gw_ref,) = torch.autograd.grad(y, w, dy)
    xp = tc.split_planes_torch(x.float().permute(0, 2, 3, 1).contiguous())
    dp = tc.split_planes_torch(dy.float().permute(0, 2, 3, 1).contiguous())
    gw = tc.conv_wgrad(dp, cout, xp, cin, k, k)
    torch.cuda.synchronize()
    err = ((gw.double() - gw_ref).abs().max() / gw_ref.abs().max()).item()
    wf = w.detach().float().requires_grad_(True)
    (gt,) = torch.autograd.grad(F.conv2d(x.float(), wf, None, padding=k // 2), wf, dy.float())
    err_t = ((gt.double() - gw_ref).abs().max() / gw_ref.abs().max()).item()
    print(f"B{B} {H}x{W} {cin}->{cout} k{k}: wgrad rel err {err:.2e} (torch fp32 {err_t:.2e})", flush=True)
    return err

worst = 0
for c in [(2, 8, 8, 64, 64, 1), (2, 8, 8, 64, 128, 3), (4, 24, 24, 128, 256, 3), (2, 24, 24, 512, 512, 3), (1, 1, 4096, 320, 384, 1),
          (1, 1, 1000, 67, 32, 1), (3, 16, 16, 128, 18, 1), (2, 48, 48, 256, 64, 3), (1, 1, 512, 3, 16, 1)]:
    worst = max(worst, check(*c))
print("worst", worst)

def bench(B, H, W, cin, cout, k, n=10):
    x = torch.randn(B, H, W, cin, device=dev); dy = torch.randn(B, H, W, cout, device=dev)
    xp = tc.split_planes_torch(x); dp = tc.split_planes_torch(dy)
    for _ in range(3): tc.conv_wgrad(dp, cout, xp, cin, k, k)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): tc.conv_wgrad(dp, cout, xp, cin, k, k)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    fl = 2.0 * B * H * W * cin * cout * k * k
    print(f"time wgrad B{B} {H}x{W} {cin}->{cout} k{k}: {ms:.3f} ms  {fl/ms/1e9:.1f} TFLOP/s useful", flush=True)
bench(32, 48, 48, 1024, 256, 3)
bench(32, 24, 24, 512, 512, 3)
bench(32, 24, 24, 2560, 1024, 1)
bench(32, 192, 192, 64, 64, 3)
bench(32, 192, 192, 64, 128, 1)
bench(1, 1, 32768, 512, 512, 1)
