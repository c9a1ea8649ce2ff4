"""Times the tcgen05 GEMM on per-point-MLP shapes (rows x K -> N, bias + ReLU + operand-plane output) under the current
environment knobs (ISTNET_BK, ISTNET_CG_SMEM_KB, ISTNET_CG_THREADS): python tools/time_rows_gemm.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from istnet_b200 import nhwc as K, tc
from istnet_b200.nhwc import Act

shapes = [(32768, 512, 512), (32768, 256, 256), (32768, 320, 384), (32768, 128, 256), (65536, 128, 128), (262144, 32, 64)]
for rows, cin, cout in shapes:
    x = torch.randn(1, 1, rows, cin, device="cuda")
    w = torch.randn(1, cout, cin, device="cuda") * 0.05
    ap, wp = tc.split_planes_torch(x), tc.split_planes_torch(w)
    bias = torch.zeros(cout, device="cuda")
    opl = K.empty_planes(1, 1, rows, cout, "cuda")
    a = Act(1, 1, rows, cin, None, ap)
    f = lambda: K.conv_gemm(a, wp, cout, 1, 1, bias=bias, relu=True, out_pl=opl)
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        f()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1000
    print(f"rows={rows} K={cin} N={cout}: {us:7.1f} us  {2.0 * rows * cin * cout * 6 / us / 1e6:7.1f} MMA-TFLOP/s")
