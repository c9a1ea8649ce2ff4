import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from istnet_b200 import tc
def bench(B, H, W, cin, cout, k, ns=2, n=10):
    x = torch.randn(B, H, W, cin, device="cuda"); dy = torch.randn(B, H, W, cout, device="cuda")
    xp = tc.split_planes_torch(x, ns); dp = tc.split_planes_torch(dy, ns)
    for _ in range(3): tc.conv_wgrad(dp, cout, xp, cin, k, k)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): tc.conv_wgrad(dp, cout, xp, cin, k, k)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    fl = 2.0 * B * H * W * cin * cout * k * k
    print(f"  wgrad B{B} {H}x{W} {cin}->{cout} k{k}: {ms:.3f} ms  {fl/ms/1e9:.1f} TFLOP/s", flush=True)
bench(32, 192, 192, 64, 64, 3); bench(32, 96, 96, 256, 64, 3); bench(32, 48, 48, 64, 64, 3); bench(32, 192, 192, 64, 128, 1)
bench(1, 1, 524288, 16, 32, 1); bench(1, 1, 131072, 67, 32, 1); bench(1, 1, 32768, 320, 384, 1); bench(32, 96, 96, 64, 256, 3); bench(32, 24, 24, 512, 512, 3)
