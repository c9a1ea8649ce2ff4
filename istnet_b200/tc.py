"""Scratch/test front-end of the tcgen05 dense kernels on plain tensors (the product path uses nhwc.py / rows_engine.py)."""
import torch

from . import nhwc as K
from .nhwc import Act


def split_planes_torch(x, nsplit=None):
    """Reference splitter (torch ops): x (..., C) FP32 -> bf16 planes [nsplit, ..., pad8(C)]."""
    nsplit = nsplit or K.NSPLIT
    c = x.shape[-1]
    cs = K.pad8(c)
    out = torch.zeros(nsplit, *x.shape[:-1], cs, dtype=torch.bfloat16, device=x.device)
    r = x.clone()
    for i in range(nsplit):
        h = r.to(torch.bfloat16)
        out[i][..., :c] = h
        r = r - h.float()
    return out


def conv_gemm(act_pl, cin, wgt_pl, cout, kh, kw, bias=None, relu=False, out_split=False):
    """act_pl: [ns,B,H,W,cs]; wgt_pl: [ns,taps,cout,cs] -> (out FP32 [B,H,W,cout], out planes | None)"""
    _, B, H, W, _ = act_pl.shape
    x = Act(B, H, W, cin, None, act_pl)
    out = torch.empty(B, H, W, cout, dtype=torch.float32, device=act_pl.device)
    opl = K.empty_planes(B, H, W, cout, act_pl.device) if out_split else None
    K.conv_gemm(x, wgt_pl, cout, kh, kw, bias=bias, relu=relu, out_f32=out, out_pl=opl)
    return out, opl


def conv_wgrad(dy_pl, cout, x_pl, cin, kh, kw):
    _, B, H, W, _ = x_pl.shape
    return K.conv_wgrad(dy_pl, cout, Act(B, H, W, cin, None, x_pl), kh, kw)
