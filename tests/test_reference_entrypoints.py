"""the reference's own entry-point plumbing — `gorilla.Config.fromfile` on its YAML files, its `Solver`
(utils/solver.py: Adam + CyclicLR + BNMomentumScheduler loop, log buffer, TensorBoard writer) and the checkpoint round trip —
executes against the shims under compat/ (gorilla-core / tensorboardX / matplotlib are not installable offline).

Runs where /root/reference exists (the build container); the GPU box has no reference tree.  The product models need CUDA (no CPU
fallback by design), so the loop is driven with a small stand-in nn.Module that honours the IST_Net dict-in / dict-out surface and
contains BatchNorm layers; what is under test here is the solver / config / checkpoint plumbing around the hot path, not the path."""
import os
import sys

import pytest
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "utils")), reason="reference tree not present")


@pytest.fixture()
def ref_paths(monkeypatch, tmp_path):
    saved = list(sys.path)
    sys.path.insert(0, os.path.join(ROOT, "compat"))  # gorilla / tensorboardX / matplotlib shims + flat model imports
    for d in ("utils", "provider", "model", os.path.join("model", "pointnet2")):
        sys.path.append(os.path.join(REF, d))  # as train.py:11-14
    monkeypatch.chdir(tmp_path)
    yield
    sys.path[:] = saved
    for k in [k for k in sys.modules if k.split(".")[0] in ("gorilla", "tensorboardX", "matplotlib", "solver", "scheduler", "evaluation_utils",
                                                                "vis_utils", "common_utils", "ist_net", "posenet_gt", "modules")]:
        sys.modules.pop(k, None)


def test_config_fromfile_reads_the_reference_yaml_files(ref_paths):
    import gorilla

    for name, arch in (("ist_net_default", "ist_net"), ("ist_net_freeze_world_enhancer", "ist_net"), ("posenet_gt_default", "posenet_gt")):
        cfg = gorilla.Config.fromfile(os.path.join(REF, "config", name + ".yaml"))
        assert cfg.model_arch == arch and cfg.num_category == 6
        assert cfg.train_dataset.img_size == 192 and cfg.train_dataset.sample_num == 1024
        assert cfg.bn.bn_momentum == 0.9 and cfg.bn.decay_step == 4000
        assert cfg.get("no_such_key", 7) == 7
        cfg.log_dir = "log/x"  # train.py:52-55 assigns attributes
        assert cfg["log_dir"] == "log/x"
    assert gorilla.Config.fromfile(os.path.join(REF, "config", "ist_net_freeze_world_enhancer.yaml")).freeze_world_enhancer is True


class _StandIn(nn.Module):
    """IST_Net's surface (ist_net.py:22-76) on a few CPU layers."""

    def __init__(self):
        super().__init__()
        self.f = nn.Sequential(nn.Conv1d(3, 16, 1), nn.BatchNorm1d(16), nn.ReLU(), nn.Conv1d(16, 3, 1))
        self.bn2d = nn.BatchNorm2d(3)
        self.head = nn.Linear(3, 12)
        self.feat = nn.Conv1d(3, 8, 1)

    def forward(self, d):
        pts = d["pts"]
        q = self.f(pts.transpose(1, 2)).transpose(1, 2)
        g = self.bn2d(d["rgb"]).mean((2, 3))
        h = self.head(q.mean(1) + g)
        r = h[:, :9].view(-1, 3, 3)
        fl = self.feat(q.transpose(1, 2))
        ep = {"pred_qo": q, "pts_w_local": fl, "pts_w_local_gt": fl.detach() + 0.1}
        for suf in ("", "_aux_cam", "_aux_world"):
            ep["pred_rotation" + suf], ep["pred_translation" + suf], ep["pred_size" + suf] = r, h[:, 9:12], h[:, 9:12].abs()
        return ep


class _SynthSet(torch.utils.data.Dataset):
    def __init__(self, n, seed):
        from istnet_b200.synth import make_batch

        self.d = make_batch(n, 64, 16, seed=seed)
        self.d["model"] = torch.zeros(n, 8, 3)
        self.d["sym_info"] = torch.zeros(n, 4, dtype=torch.int64)
        self.n, self.resets = n, 0

    def reset(self):  # provider/dataset.py:116-122, called at the start of every epoch (utils/solver.py:80-81)
        self.resets += 1

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        return {k: v[i] for k, v in self.d.items()}


def test_reference_solver_runs_two_iterations_and_checkpoints(ref_paths, monkeypatch, tmp_path):
    import gorilla
    from solver import Solver, get_logger  # the reference's utils/solver.py, unmodified

    from istnet_b200.model import SupervisedLoss

    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)  # no GPU in this container (solver.py:158-161)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)   # solver.py:153
    cfg = gorilla.Config.fromfile(os.path.join(REF, "config", "ist_net_default.yaml"))
    cfg.log_dir = str(tmp_path / "log")
    os.makedirs(cfg.log_dir)
    cfg.per_write = 1
    logger = get_logger(level_print=40, level_save=40, path_file=os.path.join(cfg.log_dir, "t.log"))
    model = _StandIn()
    loaders = {"syn": torch.utils.data.DataLoader(_SynthSet(6, 1), batch_size=3), "real": torch.utils.data.DataLoader(_SynthSet(4, 2), batch_size=2)}
    trainer = Solver(model=model, data_mode="Camera+Real", loss={"syn": SupervisedLoss(cfg), "real": SupervisedLoss(cfg)}, dataloaders=loaders,
                     logger=logger, cfg=cfg, start_epoch=1, start_iter=0)
    w0 = model.head.weight.detach().clone()
    info = trainer.train()
    assert trainer.iter == 2 and loaders["syn"].dataset.resets == 1
    assert set(info) >= {"loss_all", "loss_syn", "loss_real", "lr"} and all(v == v for v in info.values())
    assert not torch.equal(model.head.weight, w0)  # Adam stepped
    # BNMomentumScheduler (scheduler.py:277-303) reached every BatchNorm flavour with the schedule's start value (config: 0.9)
    assert all(abs(m.momentum - 0.9) < 1e-12 for m in model.modules() if isinstance(m, (nn.BatchNorm1d, nn.BatchNorm2d)))
    # checkpoint layout of utils/solver.py:64-68 and the resume path of train.py:87-96
    path = os.path.join(cfg.log_dir, "epoch_5.pth")
    gorilla.solver.save_checkpoint(model=model, filename=path, optimizer=trainer.optimizer, meta={"iter": trainer.iter, "epoch": 5})
    m2 = _StandIn()
    ck = gorilla.solver.load_checkpoint(model=m2, filename=path)
    assert ck["meta"] == {"iter": 2, "epoch": 5} and set(ck) >= {"model", "optimizer", "meta"}
    for a, b in zip(model.state_dict().values(), m2.state_dict().values()):
        assert torch.equal(a, b)
    assert sum(gorilla.parameter_count(model).values()) == sum(p.numel() for p in model.parameters())
