"""Micro-benchmark of the in-kernel (ticket) finish of per-channel reductions against the separate finalize launch it replaced.
usage (GPU box): python tools/bench_ticket.py"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from istnet_b200 import _C  # noqa: E402
from istnet_b200 import nhwc as K  # noqa: E402
from istnet_b200._C import c_int, c_ll, c_void_p, ptr  # noqa: E402


def timeit(fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1000.0


dev = torch.device("cuda")
for P, C in ((73728, 64), (18432, 512), (1179648, 64), (294912, 64), (32768, 256)):
    dz = torch.randn(P, C, device=dev)
    y = torch.randn(P, C, device=dev)
    zh = torch.randn(P, C, device=dev).bfloat16()
    mean, invstd, gamma, beta = (torch.rand(C, device=dev) + 0.5 for _ in range(4))
    st = K.BnState(mean, invstd, gamma, beta, batch=True)
    dy = K.empty_planes(1, 1, P, C, dev, nsplit=2)

    def run(with_tickets):
        ws = torch.empty(3 * C, dtype=torch.float64, device=dev)
        sg = torch.empty(C, device=dev)
        sgx = torch.empty(C, device=dev)
        part = torch.empty(_C.lib().istnet_reduce_ws_floats(c_ll(P), C, 3), dtype=torch.float32, device=dev)
        _C.call("bn_act_bwd", ptr(dz), K.NULL, ptr(y), c_ll(P), c_int(C), c_ll(1), ptr(mean), ptr(invstd), ptr(gamma), ptr(beta), c_int(1), K.NULL,
                ptr(zh), c_int(C), K.NULL, c_int(1), K.NULL, c_int(0), ptr(part), ptr(ws), *K._pl_args(dy), c_int(dy.shape[-1]), K.NULL, K.NULL,
                ptr(sg), ptr(sgx), _C.tickets(dev) if with_tickets else c_void_p(0))

    print(f"bn_act_bwd P={P} C={C}: ticket {timeit(lambda: run(True)):.1f} us   finalize-launch {timeit(lambda: run(False)):.1f} us")
    if hasattr(_C.lib(), "istnet_ticket_debug"):
        run(True)
        torch.cuda.synchronize()
        buf = (ctypes.c_ulonglong * 8)()
        _C.lib().istnet_ticket_debug(buf)
        t = [int(v) for v in buf]
        print("   tail stamps (us after the last CTA entered the tail): " + ", ".join(f"{(t[i] - t[0]) / 1000.0:.2f}" for i in range(1, 6)))
