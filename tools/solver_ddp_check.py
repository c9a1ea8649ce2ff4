"""Multi-GPU sanity of istnet_b200.solver.Solver (one process per GPU, launched with torch.distributed.run): every rank trains on its own
shard; after a few iterations the parameters must be IDENTICAL on all ranks (same all-reduced gradients, same Adam state) and differ
from the start.  usage: python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/solver_ddp_check.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from istnet_b200 import model as M  # noqa: E402
from istnet_b200.solver import Solver  # noqa: E402
from istnet_b200.synth import make_batch  # noqa: E402


class Cfg(dict):
    def __getattr__(self, k):
        v = self[k]
        return Cfg(v) if isinstance(v, dict) else v


rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
cfg = Cfg(max_epoch=2, num_mini_batch_per_epoch=6, per_write=2, per_val=10, log_dir="/tmp", optimizer={"lr": 0.01, "weight_decay": 0.0},
          bn={"bn_momentum": 0.9, "bn_decay": 0.5, "decay_step": 2, "bnm_clip": 0.01}, loss={"gamma1": 1.0, "gamma2": 10.0},
          freeze_world_enhancer=False)
torch.manual_seed(100 + rank)  # different initial weights per rank: the solver must broadcast rank 0's
m = M.IST_Net(6, False).to(dev).train()
n_it = 3
syn = [make_batch(3, 256, 64, seed=1000 * rank + i) for i in range(n_it)]
real = [make_batch(1, 256, 64, seed=1000 * rank + 500 + i) for i in range(n_it)]
loss_fn = M.SupervisedLoss(M.LossCfg())
sol = Solver(m, "Camera+Real", {"syn": loss_fn, "real": loss_fn}, {"syn": syn, "real": real}, None, cfg)
p_first = None
info = sol.train()
flat = torch.cat([p.detach().reshape(-1) for p in m.parameters()])
ref = flat.clone()
dist.broadcast(ref, src=0)
same = bool((flat == ref).all().item())
t = torch.tensor([1.0 if same else 0.0], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MIN)
losses = torch.tensor([info["loss_all"]], device=dev)
gathered = [torch.zeros_like(losses) for _ in range(world)]
dist.all_gather(gathered, losses)
if rank == 0:
    print(f"world {world}: parameters identical on all ranks after {n_it} iterations: {bool(t.item())}; "
          f"per-rank mean losses {[round(float(g.item()), 4) for g in gathered]}; optimizer steps {int(sol.optimizer.step_dev.item())}")
    assert bool(t.item()) and int(sol.optimizer.step_dev.item()) == n_it
dist.barrier()
dist.destroy_process_group()
