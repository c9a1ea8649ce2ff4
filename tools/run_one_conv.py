"""Launches one big image-branch convolution (default: up_1, 1024->256 3x3 @48x48, B=32) a few times — target for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from istnet_b200 import tc
B, H, W, cin, cout, k = [int(x) for x in (sys.argv[1:7] if len(sys.argv) >= 7 else (32, 48, 48, 1024, 256, 3))]
mode = sys.argv[7] if len(sys.argv) > 7 else "fwd"
ns = int(os.environ.get("ISTNET_NSPLIT", "3"))
x = torch.randn(B, H, W, cin, device="cuda"); w = torch.randn(k * k, cout, cin, device="cuda") * 0.02
ap = tc.split_planes_torch(x, ns); wp = tc.split_planes_torch(w, ns)
dy = tc.split_planes_torch(torch.randn(B, H, W, cout, device="cuda"), ns)
for _ in range(5):
    if mode == "fwd":
        tc.conv_gemm(ap, cin, wp, cout, k, k)
    elif mode == "fwdpl":  # per-point MLP layer: bias + ReLU + operand planes of the next layer, no FP32 output
        from istnet_b200 import nhwc as K
        from istnet_b200.nhwc import Act
        opl = K.empty_planes(B, H, W, cout, "cuda")
        K.conv_gemm(Act(B, H, W, cin, None, ap), wp, cout, k, k, bias=torch.zeros(cout, device="cuda"), relu=True, out_pl=opl)
    else:
        tc.conv_wgrad(dy, cout, ap, cin, k, k)
torch.cuda.synchronize()
