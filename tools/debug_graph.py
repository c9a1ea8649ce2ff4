import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from istnet_b200 import model as M
from istnet_b200.graph import GraphedTrainStep
from istnet_b200.synth import make_batch
LABELS = ("qo", "rotation_label", "translation_label", "size_label")
keys = ("rgb", "pts", "choose", "category_label", "qo")
torch.manual_seed(1)
m = M.IST_Net(6, False).cuda().train()
ones = {c: torch.ones(4, c, 1, 1, device="cuda") for c in (1024, 256, 64)}
m.rgb_cam_extractor.model.dropout_noise_fn = lambda b, c, p: ones[c]
loss_fn = M.SupervisedLoss(M.LossCfg())
batches = [{k: v.cuda() for k, v in make_batch(4, 512, 96, seed=s_).items()} for s_ in (31, 32)]
def eager(batch):
    for p in m.parameters(): p.grad = None
    ep = m({k: batch[k] for k in keys}); ep.update({k: batch[k] for k in LABELS})
    loss = loss_fn(ep); loss.backward()
    return loss.item(), {k: v.detach().clone() for k, v in ep.items() if k not in LABELS}
for trial in range(3):
    print("eager", [f"{eager(b)[0]:.6f}" for b in batches], flush=True)
for env in ("ISTNET_STREAMS",):
    pass
M.USE_SIDE_STREAMS = False
print("eager serial", [f"{eager(b)[0]:.6f}" for b in batches], flush=True)
M.USE_SIDE_STREAMS = True
step = GraphedTrainStep(m, loss_fn, batches[0], keys, LABELS, warmup=2)
for rep in range(3):
    print("graph", [f"{step(b).item():.6f}" for b in batches], flush=True)
print("eager again", [f"{eager(b)[0]:.6f}" for b in batches], flush=True)
