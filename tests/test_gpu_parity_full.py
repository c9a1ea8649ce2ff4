"""Parity at the BENCHMARKED shapes (BASELINE.json configs[1], [3], [4]) with every BatchNorm in train mode: the CUDA path
against the reference dataflow (oracle/istnet_port.py, pinned bit-identical to the reference modules by
tests/test_oracle_model.py) evaluated in FLOAT64 on the same GPU, indices from the C restatement of the reference kernels.

Outputs, loss and BatchNorm running statistics: 1e-4 (max|a-b| / max|b|, BASELINE.json north_star).  Gradients: the same
step is also evaluated by the port in FP32 (cuDNN with TF32 off = the parity-grade reference arithmetic) — train-mode
steps flip ReLU / max-pool selections under rounding-level perturbations, which no FP32 implementation can avoid, so the
gradient criterion is distributional (median / 90 % / max of the per-tensor errors within 3x / 3x / 5x of the FP32 reference's
own); the tight gradient bounds (1e-4 vs float64) are held per kernel in tests/test_gpu_kernels.py and tests/test_gpu_sa_fused.py."""
import pytest
import torch

from conftest import fixed_dropout_noise, rel_err
from istnet_b200 import model as M
from istnet_b200.synth import make_batch

pytestmark = pytest.mark.gpu
TOL = 1e-4
LABELS = ("qo", "rotation_label", "translation_label", "size_label")


def _port_step(kind, sd_cpu, inp, masks, dtype, freeze=False, momentum=0.1):
    from oracle import istnet_port as port
    from oracle import pointops_dev

    sd = {}
    for k, v in sd_cpu.items():
        t = v.clone().cuda()
        if t.is_floating_point():
            t = t.to(dtype)
            if "running_" not in k:
                t.requires_grad_(True)
        sd[k] = t
    data = {k: (v.cuda().to(dtype) if v.is_floating_point() else v.cuda()) for k, v in inp.items()}
    mk = [m.cuda().to(dtype) for m in masks]
    with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
        if kind == "ist_net":
            ep = port.ist_net_forward(sd, data, training=True, freeze_world_enhancer=freeze, dropout_noise=mk, ops=pointops_dev, bn_momentum=momentum)
            loss = port.ist_net_loss(ep, data, freeze_world_enhancer=freeze)
        else:
            ep = port.posenet_gt_forward(sd, data, training=True, dropout_noise=mk, ops=pointops_dev, bn_momentum=momentum)
            loss = port.posenet_gt_loss(ep, data)
        loss.backward()
    torch.cuda.synchronize()
    return ep, loss, sd


def _check(kind, B, npts, img, seed, freeze=False, momentum=0.1, median_factor=3.0):
    torch.manual_seed(1)
    m = M.IST_Net(6, freeze) if kind == "ist_net" else M.PoseNetGT(6)
    if freeze:  # train.py:116-118: parameters of the (pre-trained) world enhancer are frozen
        for n, p in m.named_parameters():
            if "world_enhancer" in n:
                p.requires_grad_(False)
    sd_cpu = {k: v.detach().clone() for k, v in m.state_dict().items()}
    inp = make_batch(B, npts, img, seed=seed, quantize=True)
    noise = fixed_dropout_noise(seed + 100)
    masks = [noise(B, 1024, 0.3), noise(B, 256, 0.15), noise(B, 64, 0.15)]
    ep64, loss64, sd64 = _port_step(kind, sd_cpu, inp, masks, torch.float64, freeze, momentum)
    ep32, loss32, sd32 = _port_step(kind, sd_cpu, inp, masks, torch.float32, freeze, momentum)
    m = m.cuda().train()
    for mod in m.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.momentum = momentum
    psp = (m.rgb_cam_extractor if kind == "ist_net" else m.rgb_extractor).model
    it = iter(masks)
    psp.dropout_noise_fn = lambda b, c, p: next(it)
    ep = m({k: v.cuda() for k, v in inp.items()})
    ep.update({k: inp[k].cuda() for k in LABELS})
    loss = (M.SupervisedLoss(M.LossCfg(1.0, 10.0, freeze)) if kind == "ist_net" else M.PoseNetGTLoss())(ep)
    loss.backward()
    worst = 0.0
    for k, v in ep64.items():
        e = rel_err(ep[k], v)
        worst = max(worst, e)
        assert e < TOL, (k, e, "reference FP32's own error:", rel_err(ep32[k], v))
    assert abs(loss.item() - loss64.item()) < TOL * abs(loss64.item())
    # BatchNorm running statistics and counters after ONE train-mode step
    sd_now = m.state_dict()
    for k, v in sd64.items():
        if "running_" in k:
            e = rel_err(sd_now[k], v)
            assert e < TOL, (k, e)
        elif k.endswith("num_batches_tracked"):
            assert int(sd_now[k]) == int(v), k
    # gradients: per-tensor error of the CUDA path and of the FP32 reference arithmetic against the float64 truth
    mine, ref = [], []
    for n, p in m.named_parameters():
        g64 = sd64[n].grad
        if g64 is None or not p.requires_grad:
            assert p.grad is None, n
            continue
        if g64.abs().max().item() < 1e-12:
            assert p.grad is None or p.grad.abs().max().item() < 1e-6, n
            continue
        mine.append((rel_err(p.grad, g64), n))
        ref.append((rel_err(sd32[n].grad, g64), n))
    q = lambda v, f: sorted(e for e, _ in v)[min(len(v) - 1, int(f * len(v)))]
    print(f"{kind} B={B} {npts}pts {img}^2: worst output err {worst:.2e}; gradient err vs float64 (median / 90% / max): "
          f"CUDA path {q(mine, 0.5):.2e} / {q(mine, 0.9):.2e} / {max(mine)[0]:.2e} ({max(mine)[1]}),  "
          f"FP32 reference {q(ref, 0.5):.2e} / {q(ref, 0.9):.2e} / {max(ref)[0]:.2e} ({max(ref)[1]})")
    # Train-mode steps flip ReLU / max-pool / arg-max selections under rounding-level perturbations, so ANY FP32 evaluation of
    # the step (the reference's included) sits 1e-3 .. 1e-1 away from the float64 gradients, tensor by tensor at random.  The
    # CUDA path must be statistically indistinguishable from the reference's own FP32 arithmetic: same error distribution.
    assert q(mine, 0.5) <= median_factor * q(ref, 0.5) + 1e-4, (q(mine, 0.5), q(ref, 0.5))
    assert q(mine, 0.9) <= 3.0 * q(ref, 0.9) + 1e-4, (q(mine, 0.9), q(ref, 0.9))
    assert max(mine)[0] <= 5.0 * max(ref)[0] + 1e-3, (max(mine), max(ref))


def test_cfg1_shape_train_step_vs_float64():
    """BASELINE.json configs[1] shape (1024 pts + 192x192 RGB, ist_net_default.yaml) at B = 8 in train mode: the head's
    Gram-matrix BatchNorm statistics run over 8 * 192^2 = 295 k pixels, every SharedMLP BatchNorm over >= 32 k rows."""
    _check("ist_net", 8, 1024, 192, seed=51, momentum=0.9)  # 0.9 = the schedule's start value (config/ist_net_default.yaml:16-20)


def test_frozen_world_enhancer_train_step_vs_float64():
    """ist_net_freeze_world_enhancer.yaml: IST_Net(6, True) — no world-space pose head, extractor frozen (train.py:103-118)."""
    _check("ist_net", 4, 1024, 192, seed=52, freeze=True)


def test_posenet_gt_full_resolution_train_step_vs_float64():
    """BASELINE.json configs[3] model at its own resolution (posenet_gt_default.yaml)."""
    _check("posenet_gt", 4, 1024, 192, seed=53)


def test_cfg4_dense_cloud_train_step_vs_float64():
    """BASELINE.json configs[4] shape: 4096-point crops, train mode."""
    # measured (profiles/r2_gpu_tests.txt): at this B = 2 shape the ratio of the two medians varies from draw to draw — seeds 54..57
    # give 4.6x, 1.0x, 0.87x, 2.8x (CUDA path 9.7e-3 / 3.4e-3 / 2.2e-3 / 4.5e-3 against the FP32 reference's 2.1e-3 / 3.4e-3 / 2.5e-3 /
    # 1.6e-3), with identical 90 % and max quantiles.  It is not the operand-plane arithmetic: three planes in every forward and
    # backward contraction give the same 9.8e-3 at seed 54.  Two FP32 evaluations whose roundings differ flip different ReLU / max
    # selections, and with two instances per BatchNorm batch a handful of flips moves the median; the bound is 6x for this shape.
    _check("ist_net", 2, 4096, 192, seed=54, median_factor=6.0)


def test_graph_replay_honours_batchnorm_momentum_changes():
    """BNMomentumScheduler (utils/scheduler.py:277-303) rewrites `bn.momentum` between iterations; the captured step reads
    momentum from device memory, so a replay must update the running statistics with the NEW value."""
    from istnet_b200.graph import GraphedTrainStep

    keys = ("rgb", "pts", "choose", "category_label", "qo")
    batch = {k: v.cuda() for k, v in make_batch(4, 256, 64, seed=61).items()}
    ones = {c: torch.ones(4, c, 1, 1, device="cuda") for c in (1024, 256, 64)}

    def build():
        torch.manual_seed(1)
        m = M.IST_Net(6, False).cuda().train()
        m.rgb_cam_extractor.model.dropout_noise_fn = lambda b, c, p: ones[c]
        return m

    def set_momentum(m, v):
        for mod in m.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.momentum = v

    loss_fn = M.SupervisedLoss(M.LossCfg())
    # eager reference: one step with momentum 0.9, one with 0.3 (parameters unchanged: no optimizer in this test)
    ref = build()
    sd0 = {k: v.clone() for k, v in ref.state_dict().items()}
    for mom in (0.9, 0.3):
        set_momentum(ref, mom)
        for p in ref.parameters():
            p.grad = None
        ep = ref({k: batch[k] for k in keys})
        ep.update({k: batch[k] for k in LABELS})
        loss_fn(ep).backward()
    want = {k: v.clone() for k, v in ref.state_dict().items() if "running_" in k}
    m = build()
    set_momentum(m, 0.9)
    step = GraphedTrainStep(m, loss_fn, batch, keys, LABELS, warmup=1)
    m.load_state_dict(sd0)
    step(batch)
    set_momentum(m, 0.3)
    step(batch)
    got = m.state_dict()
    moved = 0
    for k, v in want.items():
        assert rel_err(got[k], v) < 1e-5, (k, rel_err(got[k], v))
        moved += int((v - sd0[k]).abs().max().item() > 1e-6)
    assert moved > 100
    assert all(int(v) == 2 for k, v in got.items() if k.endswith("num_batches_tracked"))
