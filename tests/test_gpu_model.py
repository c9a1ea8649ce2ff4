"""GPU parity of the full hot path (IST_Net / PoseNetGT forward, SupervisedLoss, backward) against the golden
vectors produced by the reference itself and against the oracle port run live on the CPU.
Tolerance: 1e-4 relative (max|a-b| / max|b|) for float tensors — BASELINE.json north_star."""
import numpy as np
import pytest
import torch

from conftest import fixed_dropout_noise, golden_inputs, load_golden, rel_err, sd_checksum
from istnet_b200 import model as M
from istnet_b200.synth import make_batch

pytestmark = pytest.mark.gpu
TOL = 1e-4
# Gradients of a train-mode network with tiny batches (B=2..4) are chaotic (ReLU / max-pool selections flip under
# rounding-level perturbations): the reference's own FP32 gradients deviate from a FLOAT64 evaluation of the same step by
# 1e-3..2e-1 per tensor (tests/golden/train_b4.npz `referr_grad_*`).  Where no float64 truth is stored, gradient norms
# are held to GRAD_TOL; outputs, losses and running statistics are always held to 1e-4.
GRAD_TOL = 3e-2
LABELS = ("qo", "rotation_label", "translation_label", "size_label")


def cuda_inputs(inp):
    return {k: v.cuda() for k, v in inp.items()}


def test_cfg0_eval_forward_matches_reference_golden():
    z = load_golden("cfg0_eval.npz")
    torch.manual_seed(1)
    m = M.IST_Net(6, False)
    assert abs(sd_checksum(m.state_dict()) - float(z["sd_checksum"])) < 1e-6 * float(z["sd_checksum"])
    m = m.cuda().eval()
    with torch.no_grad():
        ep = m(cuda_inputs(golden_inputs(z)))
    assert set(ep) == {"pred_qo", "pred_rotation", "pred_translation", "pred_size"}
    for k, v in ep.items():
        assert rel_err(v, z["out_" + k]) < TOL, (k, rel_err(v, z["out_" + k]))


def test_cfg4_dense_cloud_eval_forward_matches_reference_golden():
    """4096-point crop (BASELINE.json configs[4] shape): FPS over 4096 points, SA1 on 512 x {16,32} neighbours of a 4x denser
    surface, FP0 and every per-point MLP on 4096 rows; eval forward with non-trivial BatchNorm statistics vs the reference."""
    from conftest import perturb_batchnorm

    z = load_golden("cfg4_eval.npz")
    torch.manual_seed(1)
    m = M.IST_Net(6, False)
    perturb_batchnorm(m, seed=41)
    assert abs(sd_checksum(m.state_dict()) - float(z["sd_checksum"])) < 1e-6 * float(z["sd_checksum"])
    m = m.cuda().eval()
    with torch.no_grad():
        ep = m(cuda_inputs(golden_inputs(z)))
    for k, v in ep.items():
        assert rel_err(v, z["out_" + k]) < TOL, (k, rel_err(v, z["out_" + k]))


def _train_step(m, inp, loss_mod, noise_seed, momentum=None):
    psp = (m.rgb_cam_extractor if hasattr(m, "rgb_cam_extractor") else m.rgb_extractor).model
    psp.dropout_noise_fn = fixed_dropout_noise(noise_seed)
    if momentum is not None:
        for mod in m.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.momentum = momentum
    m.train()
    ep = m(cuda_inputs(inp))
    ep.update({k: inp[k].cuda() for k in LABELS})
    loss = loss_mod(ep)
    loss.backward()
    return ep, loss


def test_train_step_matches_reference_golden():
    """B=4 train step (fwd + SupervisedLoss + bwd) against the reference.  The golden file holds the reference's FP32
    result AND the same step evaluated by the reference modules in FLOAT64 (ground truth) with the reference's own FP32
    deviation from it per tensor (`referr_*`).  Outputs / loss / running stats: 1e-4 against the truth and the FP32
    reference.  Gradients: this tiny-batch train-mode step is chaotic (ReLU / max-pool selections flip under 1e-7
    perturbations — the reference's own FP32 gradients are up to 2e-1 off the truth), so each gradient must be within
    max(2e-3, 20 x the reference's own deviation) of the truth; the well-conditioned, tight (1e-5-level) checks of every
    backward kernel live in tests/test_gpu_kernels.py."""
    z = load_golden("train_b4.npz")
    torch.manual_seed(1)
    m = M.IST_Net(6, False).cuda()
    ep, loss = _train_step(m, golden_inputs(z), M.SupervisedLoss(M.LossCfg(1.0, 10.0, False)), 77, momentum=0.9)
    for k in z:
        if k.startswith("out_"):
            assert rel_err(ep[k[4:]], z[k]) < TOL, (k, rel_err(ep[k[4:]], z[k]))
            assert rel_err(ep[k[4:]], z["out64_" + k[4:]]) < TOL, (k, "vs float64 truth")
    assert abs(loss.item() - float(z["loss"])) < TOL * abs(float(z["loss"]))
    assert abs(loss.item() - float(z["loss64"])) < TOL * abs(float(z["loss64"]))
    params = dict(m.named_parameters())
    sd = m.state_dict()
    gmax = max(float(z[k]) for k in z if k.startswith("gradnorm64_"))
    closer, total, failures = 0, 0, []
    for k in z:
        if k.startswith("gradnorm64_"):
            n = k[11:]
            g = params[n].grad
            truth = float(z[k])
            mine = g.double().norm().item()
            if truth < 1e-7 * gmax:  # analytically-zero gradients (biases feeding a train-mode BatchNorm)
                assert mine < 1e-6 * gmax, n
                continue
            bound = max(20 * TOL, 20.0 * float(z["referr_grad_" + n]))
            if abs(mine - truth) / truth > bound:
                failures.append((n, abs(mine - truth) / truth, bound))
            if ("grad64_" + n) in z:
                e = rel_err(g, z["grad64_" + n])
                total += 1
                closer += e <= float(z["referr_grad_" + n])
                if e > bound:
                    failures.append((n, e, bound))
        elif k.startswith("stat_"):
            assert rel_err(sd[k[5:]], z[k]) < TOL, k
    assert not failures, failures[:10]
    for n in z["nograd"]:
        assert params[str(n)].grad is None, n
    print(f"gradients closer to the float64 truth than the reference's own FP32 result: {closer}/{total}")


def test_posenet_gt_matches_reference_golden():
    z = load_golden("posenet_gt_b2.npz")
    torch.manual_seed(1)
    m = M.PoseNetGT(6).cuda()
    ep, loss = _train_step(m, golden_inputs(z), M.PoseNetGTLoss(), 78)
    for k in z:
        if k.startswith("out_"):
            assert rel_err(ep[k[4:]], z[k]) < TOL, (k, rel_err(ep[k[4:]], z[k]))
    assert abs(loss.item() - float(z["loss"])) < TOL * abs(float(z["loss"]))
    params = dict(m.named_parameters())
    for n in z["nograd"]:
        assert params[str(n)].grad is None, n
    for k in z:
        if k.startswith("gradnorm_"):
            assert abs(params[k[9:]].grad.double().norm().item() - float(z[k])) <= GRAD_TOL * float(z[k]) + 1e-10, k


def test_full_resolution_train_step_matches_oracle_port():
    """B=2 at the bench resolution (1024 pts, 192x192): oracle port on the host CPU vs the CUDA path, forward + loss +
    backward through the training branch of IST_Net with the BatchNorm layers in eval() (running statistics): this removes
    the batch-statistics amplification that makes tiny-batch train-mode gradients chaotic, so gradients can be held to a
    much tighter bound than in train mode.  What remains are ReLU / max-pool selection flips between the two FP32
    evaluations (measured: median 3.5e-4, 90 % 1.4e-3, max 1.5e-2; identical with 3-plane backward contractions, i.e. not a
    precision effect of the kernels) — bounds: median 2e-3, no tensor beyond 5e-2."""
    from oracle import istnet_port as port

    torch.manual_seed(1)
    m = M.IST_Net(6, False)
    g = torch.Generator().manual_seed(3)
    for mod in m.modules():  # non-trivial running statistics / affine parameters
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.running_mean.copy_(0.1 * torch.randn(mod.num_features, generator=g))
            mod.running_var.copy_(0.5 + torch.rand(mod.num_features, generator=g))
            mod.weight.data.copy_(0.5 + torch.rand(mod.num_features, generator=g))
            mod.bias.data.copy_(0.1 * torch.randn(mod.num_features, generator=g))
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    for k, v in sd.items():
        if v.is_floating_point() and "running_" not in k:
            v.requires_grad_(True)
    inp = make_batch(2, 1024, 192, seed=21, quantize=True)
    noise = fixed_dropout_noise(5)
    masks = [noise(2, 1024, 0.3), noise(2, 256, 0.15), noise(2, 64, 0.15)]
    ep_o = port.ist_net_forward(sd, inp, training=True, dropout_noise=masks, bn_training=False)
    loss_o = port.ist_net_loss(ep_o, inp)
    loss_o.backward()
    m = m.cuda().train()
    for mod in m.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.eval()
    m.rgb_cam_extractor.model.dropout_noise_fn = fixed_dropout_noise(5)
    ep = m(cuda_inputs(inp))
    ep.update({k: inp[k].cuda() for k in LABELS})
    loss = M.SupervisedLoss(M.LossCfg())(ep)
    loss.backward()
    for k in ep_o:
        assert rel_err(ep[k], ep_o[k]) < TOL, (k, rel_err(ep[k], ep_o[k]))
    assert abs(loss.item() - loss_o.item()) < TOL * abs(loss_o.item())
    errs = []
    for n, p in m.named_parameters():
        go = sd[n].grad
        if go is None:
            assert p.grad is None, n
            continue
        errs.append((rel_err(p.grad, go), n))
    errs.sort()
    med = errs[len(errs) // 2][0]
    print(f"gradient rel err vs oracle port: median {med:.2e}, 90% {errs[int(0.9 * len(errs))][0]:.2e}, max {errs[-1][0]:.2e} ({errs[-1][1]})")
    assert med < 2e-3, med
    assert errs[-1][0] < 5e-2, errs[-5:]


def test_cuda_graph_step_equals_eager_step():
    """The whole-step CUDA graph (multi-stream capture: image branch / extractors / pose heads / weight-gradient side
    streams) must reproduce the eager step: same loss and the same gradients on repeated replays with fresh inputs."""
    from istnet_b200.graph import GraphedTrainStep

    torch.manual_seed(1)
    m = M.IST_Net(6, False).cuda().train()
    ones = {c: torch.ones(4, c, 1, 1, device="cuda") for c in (1024, 256, 64)}  # device tensors: no H2D copy inside the capture
    m.rgb_cam_extractor.model.dropout_noise_fn = lambda b, c, p: ones[c]  # same (all-pass) masks in both runs
    loss_fn = M.SupervisedLoss(M.LossCfg())
    keys = ("rgb", "pts", "choose", "category_label", "qo")
    batches = [{k: v.cuda() for k, v in make_batch(4, 512, 96, seed=s_).items()} for s_ in (31, 32)]

    def eager(batch):
        for p in m.parameters():
            p.grad = None
        ep = m({k: batch[k] for k in keys})
        ep.update({k: batch[k] for k in LABELS})
        loss = loss_fn(ep)
        loss.backward()
        return loss.item(), {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None}

    sd0 = {k: v.clone() for k, v in m.state_dict().items()}
    ref = [eager(b) for b in batches]
    m.load_state_dict(sd0)  # BN running statistics back to the start
    step = GraphedTrainStep(m, loss_fn, batches[0], keys, LABELS, warmup=2)
    m.load_state_dict(sd0)
    for rep in range(2):
        for (loss_e, grads_e), batch in zip(ref, batches):
            loss_g = step(batch).item()
            assert abs(loss_g - loss_e) <= 1e-6 * abs(loss_e), (loss_g, loss_e)  # the forward pass is bitwise reproducible
            errs = []
            for n, p in m.named_parameters():
                if n in grads_e and grads_e[n].abs().max().item() >= 1e-7:
                    errs.append(rel_err(p.grad, grads_e[n]))
            errs.sort()
            # not bit-identical: the three scatter-adds use float atomics (as in the reference) and this B=4 train-mode step
            # is chaotic (see test_train_step_matches_reference_golden); gross errors (a missed dependency in the captured
            # multi-stream graph would give O(1)) are what this guards against
            assert errs[len(errs) // 2] < 2e-3 and errs[-1] < 0.1, (errs[len(errs) // 2], errs[-1])


def test_inference_engine_buckets_match_direct_eval_forward():
    """Eval-mode inference path (utils/solver.py:217-241): graph-per-bucket engine with padded instance counts and on-device pose
    assembly vs the plain eval forward + the reference's assembly arithmetic."""
    from conftest import perturb_batchnorm
    from istnet_b200.infer import InferenceEngine

    torch.manual_seed(1)
    m = M.IST_Net(6, False)
    perturb_batchnorm(m, seed=43)
    m = m.cuda().eval()
    eng = InferenceEngine(m, npts=256, img=64)
    for b, seed in ((3, 71), (1, 72), (4, 73), (3, 74)):
        d = make_batch(b, 256, 64, seed=seed)
        inp = {"rgb": d["rgb"].cuda(), "pts": d["pts"].cuda(), "choose": d["choose"].cuda(), "category_label": d["category_label"].reshape(-1).cuda()}
        rts, scales = eng(inp)
        with torch.no_grad():
            ep = m(inp)
        s = torch.norm(ep["pred_size"], dim=1, keepdim=True)
        want = torch.eye(4, device="cuda").unsqueeze(0).repeat(b, 1, 1)
        want[:, :3, 3] = ep["pred_translation"]
        want[:, :3, :3] = ep["pred_rotation"] * s.unsqueeze(2)
        assert rel_err(rts, want) < 1e-6 and rel_err(scales, ep["pred_size"] / s) < 1e-6
    assert sorted(eng._graphs) == [1, 4]


class _Cfg(dict):
    """attribute dictionary with .get like gorilla.Config"""

    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError:
            raise AttributeError(k)
        return _Cfg(v) if isinstance(v, dict) else v


def test_stall_free_solver_loop_matches_the_reference_loop():
    """istnet_b200.solver.Solver (captured step + flat Adam, CyclicLR and BNMomentumScheduler as device scalars,
    deferred loss reads) against the loop of utils/solver.py:75-127 written out with torch.optim.Adam on the eager model: same
    learning-rate and BatchNorm-momentum sequence, same losses, same parameter updates, warm-up / capture leave no trace."""
    import copy

    from istnet_b200.solver import Solver, bn_momentum_at, cyclic_lr

    cfg = _Cfg(max_epoch=2, num_mini_batch_per_epoch=6, per_write=2, per_val=10, log_dir="/tmp",
               optimizer={"lr": 0.01, "weight_decay": 0.0}, bn={"bn_momentum": 0.9, "bn_decay": 0.5, "decay_step": 2, "bnm_clip": 0.01},
               loss={"gamma1": 1.0, "gamma2": 10.0}, freeze_world_enhancer=False)
    torch.manual_seed(3)
    m = M.IST_Net(6, False).cuda().train()
    m_ref = copy.deepcopy(m)
    ones = {c: torch.ones(4, c, 1, 1, device="cuda") for c in (1024, 256, 64)}
    for mm in (m, m_ref):
        mm.rgb_cam_extractor.model.dropout_noise_fn = lambda b, c, p: ones[c]
    n_it = 4
    syn = [make_batch(3, 256, 64, seed=100 + i) for i in range(n_it)]
    real = [make_batch(1, 256, 64, seed=200 + i) for i in range(n_it)]
    loss_fn = M.SupervisedLoss(M.LossCfg())
    p0 = {n: p.detach().clone() for n, p in m.named_parameters()}

    # the reference loop (utils/solver.py:83-99,152-188)
    opt = torch.optim.Adam(m_ref.parameters(), lr=1e-5, weight_decay=0.0)
    step_up = cfg.max_epoch * cfg.num_mini_batch_per_epoch // 6
    ref_losses, ref_lrs = [], []
    for it in range(n_it):
        lr = cyclic_lr(it, 1e-5, 1e-3, step_up)
        for g in opt.param_groups:
            g["lr"] = lr
        mom = bn_momentum_at(it, 0.9, 0.5, 2, 0.01)
        for mod in m_ref.modules():
            if isinstance(mod, torch.nn.modules.batchnorm._BatchNorm):
                mod.momentum = mom
        opt.zero_grad()
        data = {k: torch.cat([syn[it][k], real[it][k]]).cuda() for k in syn[it]}
        ep = m_ref({k: data[k] for k in ("rgb", "pts", "choose", "category_label", "qo")})
        ep.update({k: data[k] for k in LABELS})
        ls = loss_fn({k: v[:3] for k, v in ep.items()})
        lr_ = loss_fn({k: v[3:] for k, v in ep.items()})
        la = (ls * 3 + lr_ * 1) / 4
        la.backward()
        opt.step()
        ref_losses.append((la.item(), ls.item(), lr_.item()))
        ref_lrs.append(lr)
    assert ref_lrs[0] == 1e-5 and abs(ref_lrs[2] - 1e-3) < 1e-12 and ref_lrs[1] > ref_lrs[0]

    logged = []

    class _Log:
        def info(self, s):
            logged.append(s)

        warning = info

    sol = Solver(m, "Camera+Real", {"syn": loss_fn, "real": loss_fn}, {"syn": syn, "real": real}, _Log(), cfg, log_lag=1)
    hist = {}
    orig_update = sol.log_buffer.update
    sol.log_buffer.update = lambda d, count=1: (orig_update(d), [hist.setdefault(k, []).append(v) for k, v in d.items()])[0]
    info = sol.train()
    assert sol.iter == n_it and int(sol.optimizer.step_dev.item()) == n_it
    assert hist["lr"] == ref_lrs
    got = list(zip(hist["loss_all"], hist["loss_syn"], hist["loss_real"]))
    assert len(got) == n_it and any("Train - " in s for s in logged)
    for a, b in zip(got[0], ref_losses[0]):
        assert abs(a - b) <= 1e-5 * abs(b), (got[0], ref_losses[0])  # same weights, same kernels: the first step is identical
    for g_, r_ in zip(got[1:], ref_losses[1:]):
        for a, b in zip(g_, r_):
            assert abs(a - b) <= 2e-2 * abs(b), (got, ref_losses)  # later steps see the (chaotic, train-mode B=4) updates
    assert abs(info["loss_all"] - sum(r[0] for r in ref_losses) / n_it) <= 2e-2 * abs(info["loss_all"])
    # parameters moved the same way; BatchNorm buffers advanced exactly n_it times with the scheduled momentum
    errs = []
    ref_params = dict(m_ref.named_parameters())
    for n, p in m.named_parameters():
        d_ref = ref_params[n].detach() - p0[n]
        if d_ref.abs().max().item() > 0:
            errs.append(rel_err(p.detach() - p0[n], d_ref))
    errs.sort()
    # Adam's update is sign-like (m / sqrt(v): +-lr per element on the first step), so on tensors whose gradients carry the chaotic
    # train-mode noise of this B=4 step (see test_cuda_graph_step_equals_eager_step) a single flipped near-zero element already
    # gives a max-norm error of 2; the tensors with deterministic gradients (pose tails ...) must agree to FP32 rounding, which
    # they only do if every step saw the right learning rate, step count and moments.  Adam's arithmetic itself:
    # test_gpu_kernels.py::test_adam_flat_kernel_matches_torch_adam.
    exact = sum(1 for e in errs if e < 1e-3)
    print(f"solver vs reference loop: {len(errs)} tensors, update error quantiles min / 25% / median / max = "
          f"{errs[0]:.1e} / {errs[len(errs) // 4]:.1e} / {errs[len(errs) // 2]:.1e} / {errs[-1]:.1e}; {exact} tensors < 1e-3")
    assert exact >= 12 and errs[len(errs) // 2] < 1.0, (exact, errs[len(errs) // 2], errs[-1])
    assert abs(sol.optimizer.lr_dev.item() - ref_lrs[-1]) <= 1e-9  # the device scalar holds the last scheduled learning rate
    bufs_ref = dict(m_ref.named_buffers())
    for n, b in m.named_buffers():
        if n.endswith("num_batches_tracked"):
            assert int(b.item()) == int(bufs_ref[n].item()) == n_it, n
        elif n.endswith("running_mean") and bufs_ref[n].abs().max().item() > 1e-3:
            assert rel_err(b, bufs_ref[n]) < 5e-2, (n, rel_err(b, bufs_ref[n]))
    assert all(abs(mod.momentum - bn_momentum_at(n_it - 1, 0.9, 0.5, 2, 0.01)) < 1e-12 for mod in m.modules()
               if isinstance(mod, torch.nn.modules.batchnorm._BatchNorm))


def test_solver_optimizer_state_round_trip_resumes_identically():
    """Checkpoint / resume of the stall-free loop (utils/solver.py:64-68 + train.py:87-96): parameters + BatchNorm buffers through
    state_dict, Adam's flat moments and step count through optimizer_state() / load_optimizer_state(), iteration counter through
    start_iter (it drives CyclicLR and the BatchNorm momentum schedule).  Two more iterations after the resume equal the same two
    iterations of the uninterrupted run."""
    import copy

    from istnet_b200.solver import Solver

    cfg = _Cfg(max_epoch=2, num_mini_batch_per_epoch=6, per_write=100, per_val=10, log_dir="/tmp",
               optimizer={"lr": 0.01, "weight_decay": 0.0}, bn={"bn_momentum": 0.9, "bn_decay": 0.5, "decay_step": 2, "bnm_clip": 0.01},
               loss={"gamma1": 1.0, "gamma2": 10.0}, freeze_world_enhancer=False)
    torch.manual_seed(4)
    m = M.IST_Net(6, False).cuda().train()
    ones = {c: torch.ones(4, c, 1, 1, device="cuda") for c in (1024, 256, 64)}
    m.rgb_cam_extractor.model.dropout_noise_fn = lambda b, c, p: ones[c]
    syn = [make_batch(3, 256, 64, seed=300 + i) for i in range(4)]
    real = [make_batch(1, 256, 64, seed=400 + i) for i in range(4)]
    loss_fn = M.SupervisedLoss(M.LossCfg())
    sol = Solver(m, "Camera+Real", {"syn": loss_fn, "real": loss_fn}, {"syn": syn[:2], "real": real[:2]}, None, cfg)
    sol.train()
    ckpt_model = copy.deepcopy(m.state_dict())
    ckpt_opt = sol.optimizer_state()
    assert ckpt_opt["step"] == 2 and sol.iter == 2
    sol.dataloaders = {"syn": syn[2:], "real": real[2:]}
    info_a = sol.train()
    want = {k: v.detach().clone() for k, v in m.state_dict().items()}

    torch.manual_seed(5)  # different initial weights: everything must come from the checkpoint
    m2 = M.IST_Net(6, False).cuda().train()
    m2.rgb_cam_extractor.model.dropout_noise_fn = lambda b, c, p: ones[c]
    m2.load_state_dict(ckpt_model)
    sol2 = Solver(m2, "Camera+Real", {"syn": loss_fn, "real": loss_fn}, {"syn": syn[2:], "real": real[2:]}, None, cfg, start_iter=2)
    sol2.load_optimizer_state(ckpt_opt)
    info_b = sol2.train()
    assert sol2.iter == 4 and int(sol2.optimizer.step_dev.item()) == 4
    assert abs(info_a["lr"] - info_b["lr"]) < 1e-15
    assert abs(info_a["loss_all"] - info_b["loss_all"]) <= 2e-2 * abs(info_a["loss_all"])
    errs = sorted(rel_err(v, want[k]) for k, v in m2.state_dict().items() if v.is_floating_point() and want[k].abs().max().item() > 0)
    exact = sum(1 for e in errs if e < 1e-4)
    # same caveat as test_stall_free_solver_loop_matches_the_reference_loop: float atomics make two runs of this chaotic B=4 step differ,
    # and Adam's sign-like update magnifies it on near-zero gradient elements; most tensors must nevertheless coincide
    print(f"solver resume: {len(errs)} tensors, {exact} within 1e-4, median {errs[len(errs) // 2]:.1e}, max {errs[-1]:.1e}")
    assert exact >= len(errs) // 4 and errs[len(errs) // 2] < 5e-2, (exact, len(errs), errs[len(errs) // 2], errs[-1])
    for k, v in m2.state_dict().items():
        if k.endswith("num_batches_tracked"):
            assert int(v.item()) == int(want[k].item()) == 4, k
