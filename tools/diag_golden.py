"""Prints the per-tensor deviations of the CUDA path from the reference goldens (outputs, gradients, running stats)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from conftest import fixed_dropout_noise, golden_inputs, load_golden, rel_err
from istnet_b200 import model as M
LABELS = ("qo", "rotation_label", "translation_label", "size_label")
z = load_golden("train_b4.npz")
torch.manual_seed(1)
m = M.IST_Net(6, False).cuda()
m.rgb_cam_extractor.model.dropout_noise_fn = fixed_dropout_noise(77)
for mod in m.modules():
    if isinstance(mod, torch.nn.BatchNorm2d):
        mod.momentum = 0.9
m.train()
inp = golden_inputs(z)
ep = m({k: v.cuda() for k, v in inp.items()})
ep.update({k: inp[k].cuda() for k in LABELS})
loss = M.SupervisedLoss(M.LossCfg())(ep)
loss.backward()
print("loss", loss.item(), float(z["loss"]), abs(loss.item() - float(z["loss"])) / float(z["loss"]))
for k in z:
    if k.startswith("out_"):
        print(f"  {k:32s} {rel_err(ep[k[4:]], z[k]):.2e}")
params = dict(m.named_parameters())
rows = []
for k in z:
    if k.startswith("gradnorm_"):
        g = params[k[9:]].grad
        if g is None:
            rows.append((float("inf"), k[9:], 0, float(z[k]))); continue
        rows.append((abs(g.double().norm().item() - float(z[k])) / max(float(z[k]), 1e-30), k[9:], g.double().norm().item(), float(z[k])))
rows.sort(reverse=True)
for r in rows[:25]:
    print(f"  gradnorm {r[1]:60s} rel {r[0]:.2e}  mine {r[2]:.4e} ref {r[3]:.4e}")
rows = [(rel_err(params[k[5:]].grad, z[k]), k[5:]) for k in z if k.startswith("grad_")]
rows.sort(reverse=True)
for r in rows[:15]:
    print(f"  grad {r[1]:60s} rel {r[0]:.2e}")
