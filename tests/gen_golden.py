"""Generates tests/golden/*.npz by running the UNMODIFIED reference Python (oracle/ref_harness.py) on CPU.

Run in the build container only (needs /root/reference):  python tests/gen_golden.py
Weights: the istnet_b200 modules are constructed under torch.manual_seed(1) (config rd_seed,
ist_net_default.yaml:60) and their state_dict is loaded into the reference modules — this also proves
state-dict compatibility.  Inputs: istnet_b200.synth.make_batch (seeded, quantised so that the centring
`pts - mean` is exact).  The reference's CUDA-only `_ext` is replaced by oracle/pointops_ref.c.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from istnet_b200 import model as M  # noqa: E402
from istnet_b200.synth import make_batch  # noqa: E402
from oracle import ref_harness  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
LABELS = ("qo", "rotation_label", "translation_label", "size_label")


def sd_checksum(sd):
    return float(sum(v.double().abs().sum().item() for v in sd.values() if v.is_floating_point()))


def fixed_dropout_noise(seed):
    """Deterministic (B,C,1,1) dropout masks shared by both sides (modules.py:56,62 p=0.3/0.15/0.15)."""
    g = torch.Generator().manual_seed(seed)

    def fn(b, c, p):
        return torch.empty(b, c, 1, 1).bernoulli_(1 - p, generator=g).div_(1 - p)

    return fn


def grads_summary(model):
    out = {}
    for n, p in model.named_parameters():
        if p.grad is not None:
            out[n] = p.grad.detach()
    return out


def main():
    ns = ref_harness.load()
    torch.set_num_threads(8)
    # ---------------- cfg0: single instance, 256 pts + 64x64, eval forward (BASELINE.json configs[0])
    torch.manual_seed(1)
    mine = M.IST_Net(6, False)
    ref = ns.ist_net.IST_Net(6, False)
    ref.load_state_dict(mine.state_dict())
    ref.eval()
    inp = make_batch(1, 256, 64, seed=11, quantize=True)
    with torch.no_grad():
        ep = ref({k: v.clone() for k, v in inp.items()})
    np.savez_compressed(
        os.path.join(OUT, "cfg0_eval.npz"),
        sd_checksum=sd_checksum(mine.state_dict()),
        **{"in_" + k: v.numpy() for k, v in inp.items()},
        **{"out_" + k: v.numpy() for k, v in ep.items()},
    )
    # ---------------- small train step: B=4, 256 pts (with duplicates), 64x64, fwd + SupervisedLoss + bwd
    for name, freeze in (("train_b4", False),):
        torch.manual_seed(1)
        mine = M.IST_Net(6, freeze)
        ref = ns.ist_net.IST_Net(6, freeze)
        ref.load_state_dict(mine.state_dict())
        ref.train()
        for m in ref.modules():  # BN momentum as the scheduler sets it at iteration 0 (solver.py:48-49)
            if isinstance(m, torch.nn.BatchNorm2d):
                m.momentum = 0.9
        inp = make_batch(4, 256, 64, seed=12, quantize=True, duplicates=True)
        noise = fixed_dropout_noise(77)
        masks = []
        psp = ref.rgb_cam_extractor.model

        class _Drop(torch.nn.Module):  # replaces nn.Dropout2d with the same x*noise arithmetic, fixed masks
            def __init__(self, p):
                super().__init__()
                self.p = p

            def forward(self, x):
                m = noise(x.shape[0], x.shape[1], self.p)
                masks.append(m)
                return x * m

        psp.drop_1, psp.drop_2 = _Drop(0.3), _Drop(0.15)
        ep = ref({k: v.clone() for k, v in inp.items()})
        ep.update({k: inp[k] for k in LABELS})
        loss = ns.ist_net.SupervisedLoss(M.LossCfg(1.0, 10.0, freeze))(ep)
        loss.backward()
        g = grads_summary(ref)
        sd_after = ref.state_dict()
        keep = [n for n in g if g[n].numel() <= 4096]  # full small grads; norms for everything
        # ---- FLOAT64 ground truth of the same step (reference modules in double, same indices): separates the
        # reference's own FP32 rounding noise from implementation differences
        import sys as _sys
        from oracle import pointops_any
        _sys.modules["pointnet2._ext"] = pointops_any
        ns.pointnet2_utils._ext = pointops_any
        torch.manual_seed(1)
        mine64 = M.IST_Net(6, freeze)
        ref64 = ns.ist_net.IST_Net(6, freeze)
        ref64.load_state_dict(mine64.state_dict())
        ref64 = ref64.double().train()
        for m in ref64.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.momentum = 0.9
        noise64 = fixed_dropout_noise(77)

        class _Drop64(torch.nn.Module):
            def __init__(self, p):
                super().__init__()
                self.p = p

            def forward(self, x):
                return x * noise64(x.shape[0], x.shape[1], self.p).double()

        ref64.rgb_cam_extractor.model.drop_1, ref64.rgb_cam_extractor.model.drop_2 = _Drop64(0.3), _Drop64(0.15)
        inp64 = {k: (v.double() if v.is_floating_point() else v.clone()) for k, v in inp.items()}
        ep64 = ref64(inp64)
        ep64.update({k: inp64[k] for k in LABELS})
        loss64 = ns.ist_net.SupervisedLoss(M.LossCfg(1.0, 10.0, freeze))(ep64)
        loss64.backward()
        g64 = grads_summary(ref64)
        ns.pointnet2_utils._ext = ns.pointnet2_utils._ext.__class__ and __import__("oracle.pointops", fromlist=["x"])
        _sys.modules["pointnet2._ext"] = ns.pointnet2_utils._ext
        truth = {"loss64": loss64.item()}
        truth.update({"out64_" + k: v.detach().numpy() for k, v in ep64.items() if k not in LABELS})
        truth.update({"grad64_" + n: g64[n].numpy() for n in keep})
        truth.update({"gradnorm64_" + n: np.float64(v.norm().item()) for n, v in g64.items()})
        # per-tensor deviation of the reference's own FP32 result from the truth (max|a-b|/max|b|)
        truth.update({"referr_grad_" + n: np.float64(((g[n].double() - g64[n]).abs().max() / g64[n].abs().max().clamp_min(1e-300)).item()) for n in g})
        truth.update({"referr_out_" + k: np.float64(((ep[k].detach().double() - ep64[k].detach()).abs().max() / ep64[k].detach().abs().max()).item())
                      for k in ep64 if k not in LABELS})
        np.savez_compressed(
            os.path.join(OUT, f"{name}.npz"),
            sd_checksum=sd_checksum(mine.state_dict()),
            loss=loss.item(),
            **{"in_" + k: v.numpy() for k, v in inp.items()},
            **{"out_" + k: v.detach().numpy() for k, v in ep.items() if k not in LABELS},
            **{"gradnorm_" + n: np.float64(v.double().norm().item()) for n, v in g.items()},
            **{"grad_" + n: g[n].numpy() for n in keep},
            **{"stat_" + k: v.numpy() for k, v in sd_after.items() if "running_" in k and v.numel() <= 64},
            nograd=np.array([n for n, p in ref.named_parameters() if p.grad is None]),
            **truth,
        )
    # ---------------- PoseNetGT (cfg3 shape family): B=2, 256 pts, 64x64 train step
    torch.manual_seed(1)
    mine = M.PoseNetGT(6)
    ref = ns.posenet_gt.PoseNetGT(6)
    ref.load_state_dict(mine.state_dict())
    ref.train()
    noise = fixed_dropout_noise(78)
    psp = ref.rgb_extractor.model

    class _Drop2(torch.nn.Module):
        def __init__(self, p):
            super().__init__()
            self.p = p

        def forward(self, x):
            return x * noise(x.shape[0], x.shape[1], self.p)

    psp.drop_1, psp.drop_2 = _Drop2(0.3), _Drop2(0.15)
    inp = make_batch(2, 256, 64, seed=13, quantize=True)
    ep = ref({k: v.clone() for k, v in inp.items()})
    ep.update({k: inp[k] for k in LABELS})
    loss = ns.posenet_gt.SupervisedLoss(M.LossCfg())(ep)
    loss.backward()
    g = grads_summary(ref)
    np.savez_compressed(
        os.path.join(OUT, "posenet_gt_b2.npz"),
        sd_checksum=sd_checksum(mine.state_dict()),
        loss=loss.item(),
        **{"in_" + k: v.numpy() for k, v in inp.items()},
        **{"out_" + k: v.detach().numpy() for k, v in ep.items() if k not in LABELS},
        **{"gradnorm_" + n: np.float64(v.double().norm().item()) for n, v in g.items()},
        nograd=np.array([n for n, p in ref.named_parameters() if p.grad is None]),
    )
    print("golden written:", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
