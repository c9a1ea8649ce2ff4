"""Race hunt at the ConvUnit level: forward+backward repeated with identical inputs must be bit-identical."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from istnet_b200 import nhwc as K
from istnet_b200.nhwc import Act, ConvUnit
torch.manual_seed(0)
def run(B, H, W, cin, cout, k, bn, act, n=200):
    conv = torch.nn.Conv2d(cin, cout, k, padding=k // 2, bias=True).cuda()
    bnm = torch.nn.BatchNorm2d(cout).cuda() if bn else None
    prelu = torch.nn.PReLU().cuda() if act == 2 else None
    x = torch.randn(B, H, W, cin, device="cuda"); dz = torch.randn(B, H, W, cout, device="cuda")
    unit = ConvUnit(conv.weight, conv.bias, bnm, act, prelu=prelu.weight if prelu else None, k=k)
    ref = None; bad = {}
    junk = torch.empty(32 << 20, device="cuda")
    for i in range(n):
        xin = Act(B, H, W, cin, x); xin.pl = K.empty_planes(B, H, W, cin, "cuda"); K.split(x, B * H * W, cin, xin.pl)
        out, rec = unit.forward(xin, True, True, want_f32=True)
        grads = {}
        dx, _ = unit.backward(rec, dz, None, need_dx=True, grads=grads)
        cur = {"z": out.f32, "dx": dx, "dw": grads[id(conv.weight)], "db": grads[id(conv.bias)]}
        if i % 5 == 0: junk.normal_()
        if ref is None: ref = {k_: v.clone() for k_, v in cur.items()}
        else:
            for k_, v in cur.items():
                if not torch.equal(v, ref[k_]):
                    bad[k_] = bad.get(k_, 0) + 1
                    if bad[k_] == 1: print("   first mismatch", k_, (v - ref[k_]).abs().max().item(), "ref max", ref[k_].abs().max().item(), "iter", i)
    print(f"unit B{B} {H}x{W} {cin}->{cout} k{k} bn{int(bn)} act{act}: mismatches {bad}", flush=True)
run(2, 1, 2048, 320, 256, 1, False, 2)
run(2, 1, 2048, 320, 256, 1, False, 1)
run(4, 16, 16, 64, 64, 3, True, 2)
run(4, 16, 16, 64, 128, 3, True, 1)
