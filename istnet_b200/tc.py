"""Python front-end of the tcgen05 dense kernels (include/istnet_b200.h §3).

Tensors here are channels-last "pixel matrices": act[b,h,w,c] (images) or act[rows,c] (point sets, h=b=1), carried
as bf16 (hi, lo) pairs with the channel stride padded to a multiple of 8.
"""
import torch

from . import _C
from ._C import c_int, c_void_p, ptr

NULL = c_void_p(0)


def pad8(c):
    return (c + 7) // 8 * 8


def split_bf16_torch(x, cs=None):
    """Reference splitter (torch ops): x (..., C) fp32 -> (hi, lo) bf16 (..., cs).  Used by tests; the product path
    uses the fused CUDA splitters."""
    c = x.shape[-1]
    cs = cs or pad8(c)
    hi = x.to(torch.bfloat16)
    lo = (x - hi.float()).to(torch.bfloat16)
    if cs != c:
        pad = (0, cs - c)
        hi, lo = torch.nn.functional.pad(hi, pad), torch.nn.functional.pad(lo, pad)
    return hi.contiguous(), lo.contiguous()


def pick_box(h, w):
    """Pixel tile (box_w, box_h) with box_w*box_h | 128 that wastes the least."""
    if h == 1:
        return 128, 1
    if w % 16 == 0 and h % 8 == 0:
        return 16, 8
    return 8, 8


def conv_gemm(act_hi, act_lo, cin, wgt_hi, wgt_lo, cout, kh, kw, bias=None, relu=False, out_f32=True, out_split=False, out_cs=None, split_cs=None):
    """act_{hi,lo}: bf16 [B,H,W,cs_a]; wgt_{hi,lo}: bf16 [kh*kw, cout, cs_w] -> (out fp32 [B,H,W,out_cs] | None, (hi, lo) | None)"""
    B, H, W, cs_a = act_hi.shape
    cs_w = wgt_hi.shape[-1]
    dev = act_hi.device
    out = None
    if out_f32:
        out_cs = out_cs or cout
        out = torch.empty(B, H, W, out_cs, dtype=torch.float32, device=dev)
    oh = ol = None
    if out_split:
        split_cs = split_cs or pad8(cout)
        alloc = torch.zeros if split_cs != cout else torch.empty
        oh = alloc(B, H, W, split_cs, dtype=torch.bfloat16, device=dev)
        ol = alloc(B, H, W, split_cs, dtype=torch.bfloat16, device=dev)
    bw, bh = pick_box(H, W)
    _C.call(
        "conv_gemm", ptr(act_hi), ptr(act_lo), c_int(B), c_int(H), c_int(W), c_int(cin), c_int(cs_a), ptr(wgt_hi), ptr(wgt_lo),
        c_int(cout), c_int(cs_w), c_int(kh), c_int(kw), ptr(bias) if bias is not None else NULL, c_int(1 if relu else 0),
        ptr(out) if out is not None else NULL, c_int(out_cs or 0), ptr(oh) if oh is not None else NULL,
        ptr(ol) if ol is not None else NULL, c_int(split_cs or 0), c_int(bw), c_int(bh),
    )
    return out, ((oh, ol) if out_split else None)


def conv_wgrad(dy_hi, dy_lo, cout, x_hi, x_lo, cin, kh, kw):
    """dy_{hi,lo}: bf16 [B,H,W,cs_dy]; x_{hi,lo}: bf16 [B,H,W,cs_x] -> grad_w fp32 [cout, cin, kh, kw]"""
    B, H, W, cs_dy = dy_hi.shape
    cs_x = x_hi.shape[-1]
    dev = dy_hi.device
    ks = _C.lib().istnet_wgrad_ksplit(B, H, W, cout, cin, kh, kw)
    ws = torch.empty(ks * kh * kw * cout * cin, dtype=torch.float32, device=dev)
    gw = torch.empty(cout, cin, kh, kw, dtype=torch.float32, device=dev)
    bw, bh = (64, 1) if H == 1 else (8, 8)
    _C.call(
        "conv_wgrad", ptr(dy_hi), ptr(dy_lo), c_int(cs_dy), ptr(x_hi), ptr(x_lo), c_int(cs_x), c_int(B), c_int(H), c_int(W),
        c_int(cout), c_int(cin), c_int(kh), c_int(kw), ptr(ws), c_int(ks), ptr(gw), c_int(bw), c_int(bh),
    )
    return gw
