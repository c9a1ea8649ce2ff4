import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from conftest import rel_err
from istnet_b200 import nhwc as K
from test_gpu_kernels import _unit_vs_torch
for cfg in [dict(B=2, H=1, W=2048, cin=320, cout=256, k=1, stride=1, bn=False, act=2, bias=True, noise=False),
            dict(B=1, H=1, W=2048, cin=320, cout=256, k=1, stride=1, bn=False, act=2, bias=True, noise=False),
            dict(B=1, H=1, W=2048, cin=256, cout=256, k=1, stride=1, bn=False, act=2, bias=True, noise=False),
            dict(B=1, H=1, W=2048, cin=320, cout=256, k=1, stride=1, bn=False, act=0, bias=True, noise=False),
            dict(B=1, H=1, W=2048, cin=384, cout=256, k=1, stride=1, bn=False, act=1, bias=True, noise=False),
            dict(B=1, H=1, W=2048, cin=512, cout=384, k=1, stride=1, bn=False, act=1, bias=True, noise=False)]:
    res = _unit_vs_torch(K, seed=11, **cfg)
    print(cfg["B"], cfg["cin"], cfg["cout"], cfg["act"], {n: f"{rel_err(a, b):.1e}" for n, (a, b) in res.items()})
    if cfg["cin"] == 320 and cfg["B"] == 1 and cfg["act"] == 2:
        a, b = res["dx"]
        e = (a.double() - b).abs()[0, 0]
        print("err by channel block:", [f"{e[:, i:i+32].max().item():.1e}" for i in range(0, 320, 32)])
