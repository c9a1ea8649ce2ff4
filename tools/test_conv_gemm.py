"""GPU scratch test of the tcgen05 conv/GEMM kernel against an FP64 torch reference (run under `timeout`)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from istnet_b200 import tc

torch.manual_seed(0)
dev = "cuda"

def check(B, H, W, cin, cout, k, bias, relu, split):
    x = torch.randn(B, cin, H, W, device=dev)
    w = torch.randn(cout, cin, k, k, device=dev) / (cin * k * k) ** 0.5
    bvec = torch.randn(cout, device=dev) if bias else None
    ref = F.conv2d(x.double(), w.double(), bvec.double() if bias else None, padding=k // 2)
    if relu:
        ref = ref.relu()
    ref = ref.permute(0, 2, 3, 1)
    ap = tc.split_planes_torch(x.permute(0, 2, 3, 1).contiguous())
    wp = tc.split_planes_torch(w.permute(2, 3, 0, 1).reshape(k * k, cout, cin).contiguous())
    out, sp = tc.conv_gemm(ap, cin, wp, cout, k, k, bias=bvec, relu=relu, out_split=split)
    torch.cuda.synchronize()
    err = ((out.double() - ref).abs().max() / ref.abs().max()).item()
    tf = F.conv2d(x, w, bvec, padding=k // 2)
    if relu:
        tf = tf.relu()
    err_t = ((tf.permute(0, 2, 3, 1).double() - ref).abs().max() / ref.abs().max()).item()
    msg = f"B{B} {H}x{W} {cin}->{cout} k{k} bias{int(bias)} relu{int(relu)}: rel err {err:.2e} (torch fp32 {err_t:.2e})"
    if split:
        rs = sp.float().sum(0)
        e2 = ((rs[..., :cout].double() - ref).abs().max() / ref.abs().max()).item()
        msg += f" split {e2:.2e}"
    print(msg, flush=True)
    return err

cases = [
    (2, 8, 8, 64, 64, 1, False, False, False),
    (2, 8, 8, 64, 64, 3, False, False, False),
    (2, 24, 24, 128, 256, 3, True, True, True),
    (3, 24, 24, 512, 512, 3, False, False, False),
    (1, 1, 1000, 320, 384, 1, True, True, True),
    (1, 1, 4096, 67, 32, 1, False, False, True),
    (1, 1, 300, 3, 16, 1, False, False, False),
    (2, 48, 48, 1024, 256, 3, True, False, False),
    (2, 16, 16, 128, 18, 1, True, False, False),
    (1, 1, 512, 2560, 1024, 1, True, True, False),
]
worst = 0
for c in cases:
    worst = max(worst, check(*c))
print("worst", worst)
# quick timing of the big shapes
def bench(B, H, W, cin, cout, k, n=10):
    x = torch.randn(B, H, W, cin, device=dev); w = torch.randn(k * k, cout, cin, device=dev) * 0.02
    ap = tc.split_planes_torch(x); wp = tc.split_planes_torch(w)
    for _ in range(3): tc.conv_gemm(ap, cin, wp, cout, k, k)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): tc.conv_gemm(ap, cin, wp, cout, k, k)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    fl = 2.0 * B * H * W * cin * cout * k * k
    print(f"time B{B} {H}x{W} {cin}->{cout} k{k}: {ms:.3f} ms  {fl/ms/1e9:.1f} TFLOP/s useful ({3*fl/ms/1e9:.1f} MMA)", flush=True)
bench(32, 48, 48, 1024, 256, 3)
bench(32, 24, 24, 512, 512, 3)
bench(32, 24, 24, 2560, 1024, 1)
bench(32, 192, 192, 64, 64, 3)
bench(32, 192, 192, 64, 128, 1)
bench(1, 1, 32768, 512, 512, 1)
