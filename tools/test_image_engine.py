"""GPU scratch test: image branch on the B200 engine vs the same module evaluated by PyTorch (FP64 and FP32)."""
import copy, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from istnet_b200.image import Modified_PSPNet
from istnet_b200.image_engine import image_branch
from istnet_b200.model import gather_pixels

torch.backends.cudnn.allow_tf32 = False
dev = "cuda"

def rel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()

def run(B, S, N, train):
    torch.manual_seed(1)
    net = Modified_PSPNet().to(dev)
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.momentum = 0.9
            torch.nn.init.uniform_(m.weight, 0.5, 1.5); torch.nn.init.normal_(m.bias, 0, 0.2)
    g = torch.Generator().manual_seed(3)
    masks = {}
    def noise_fn(b, c, p):
        if c not in masks:
            masks[c] = torch.empty(b, c, 1, 1).bernoulli_(1 - p, generator=g).div_(1 - p)
        return masks[c]
    net.dropout_noise_fn = noise_fn
    net.train(train)
    rgb = torch.randn(B, 3, S, S, device=dev)
    choose = torch.randint(0, S * S, (B, N), device=dev)
    d = torch.randn(B, 128, N, device=dev)
    ref64 = copy.deepcopy(net).double(); ref64.dropout_noise_fn = noise_fn
    ref32 = copy.deepcopy(net); ref32.dropout_noise_fn = noise_fn
    outs, grads, stats = {}, {}, {}
    for name, m, x in (("f64", ref64, rgb.double()), ("f32", ref32, rgb)):
        with torch.set_grad_enabled(train):
            o = gather_pixels(m(x), choose)
        outs[name] = o.detach()
        if train:
            (o * d.to(o.dtype)).sum().backward()
            grads[name] = {n: p.grad for n, p in m.named_parameters() if p.grad is not None}
        stats[name] = {n: b.clone() for n, b in m.named_buffers()}
    with torch.set_grad_enabled(train):
        o = image_branch(net, rgb, choose).transpose(1, 2)
    torch.cuda.synchronize()
    print(f"--- B{B} {S}x{S} N{N} train={train}: out rel err engine {rel(o, outs['f64']):.2e} | torch-fp32 {rel(outs['f32'], outs['f64']):.2e}")
    if train:
        (o * d).sum().backward()
        torch.cuda.synchronize()
        rows = []
        for n, p in net.named_parameters():
            if n not in grads["f64"]:
                assert p.grad is None, n
                continue
            assert p.grad is not None, n
            scale = grads["f64"][n].abs().max().item()
            rows.append((rel(p.grad, grads["f64"][n]), rel(grads["f32"][n], grads["f64"][n]), scale, n))
        rows.sort(reverse=True)
        for e, et, sc, n in rows[:12]:
            print(f"   grad {n:45s} engine {e:.2e} torch-fp32 {et:.2e} (|g|max {sc:.2e})")
        big = [r for r in rows if r[0] > 1e-3 and r[2] > 1e-6]
        print("   #params", len(rows), "worst engine", rows[0][0], "n>1e-3:", len(big))
        sb = {n: b for n, b in net.named_buffers()}
        w = max(rel(sb[n], stats["f64"][n]) for n in sb if "running" in n)
        print(f"   running stats worst rel err {w:.2e}; nbt equal: {all(torch.equal(sb[n], stats['f64'][n]) for n in sb if 'num_batches' in n)}")

run(2, 64, 256, False)
run(2, 64, 256, True)
run(4, 192, 1024, True)
