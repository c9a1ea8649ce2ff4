"""CPU checks of the host-side pieces of the path that need no kernel: the torch parts of the module surface (Ortho6d2Mat,
SupervisedLoss) against the oracle port (which is pinned to the reference modules, tests/test_oracle_model.py), and the small
planning helpers (operand-plane map, pyramid row layout, channel padding)."""
import torch

from istnet_b200 import image_engine as IE
from istnet_b200 import model as M
from istnet_b200 import nhwc
from oracle import istnet_port as port


def test_ortho6d_matches_port_and_is_orthonormal():
    g = torch.Generator().manual_seed(0)
    x, y = torch.randn(16, 3, generator=g), torch.randn(16, 3, generator=g)
    r = M.ortho6d_to_mat(x, y)
    assert torch.allclose(r, port.ortho6d_to_mat(x, y), atol=2e-7)
    eye = torch.eye(3).expand(16, 3, 3)
    assert torch.allclose(r.transpose(1, 2) @ r, eye, atol=1e-5) and torch.allclose(torch.linalg.det(r), torch.ones(16), atol=1e-5)
    # degenerate input: the reference clamps the norm at 1e-8 instead of dividing by zero (rotation_utils.py:4-9)
    z = M.ortho6d_to_mat(torch.zeros(2, 3), torch.zeros(2, 3))
    assert torch.isfinite(z).all()


def test_supervised_losses_match_port():
    g = torch.Generator().manual_seed(1)
    B, N = 3, 50
    ep = {"pred_qo": torch.randn(B, N, 3, generator=g), "qo": torch.randn(B, N, 3, generator=g) * 0.3,
          "pts_w_local": torch.randn(B, 128, N, generator=g), "pts_w_local_gt": torch.randn(B, 128, N, generator=g)}
    for suffix in ("", "_aux_cam", "_aux_world"):
        ep["pred_rotation" + suffix] = torch.randn(B, 3, 3, generator=g)
        ep["pred_translation" + suffix] = torch.randn(B, 3, generator=g)
        ep["pred_size" + suffix] = torch.rand(B, 3, generator=g)
    labels = {"rotation_label": torch.randn(B, 3, 3, generator=g), "translation_label": torch.randn(B, 3, generator=g),
              "size_label": torch.rand(B, 3, generator=g), "qo": ep["qo"]}
    ep.update(labels)
    for freeze in (False, True):
        mine = M.SupervisedLoss(M.LossCfg(1.0, 10.0, freeze))(ep)
        want = port.ist_net_loss(ep, labels, 1.0, 10.0, freeze)
        assert abs(mine.item() - want.item()) <= 1e-6 * abs(want.item())
    gt = M.PoseNetGTLoss()(ep)
    assert abs(gt.item() - port.posenet_gt_loss(ep, labels).item()) <= 1e-6 * abs(gt.item())
    # SmoothL1Dis: quadratic below the threshold, linear above (losses.py:3-22)
    a, b = torch.zeros(1, 2, 3), torch.tensor([[[0.05, 0.0, 0.0], [0.3, 0.0, 0.0]]])
    assert abs(M.SmoothL1Dis(a, b).item() - 0.5 * (0.05 ** 2 / 0.2 + (0.3 - 0.05))) < 1e-7


def test_plane_map_and_layout_helpers():
    m = IE._parse_ns_map("up_=2, layer4=2")
    assert m == {"up_": 2, "layer4": 2}
    assert IE._parse_ns_map("") == {}
    saved, IE.NS_MAP = IE.NS_MAP, m
    try:
        assert IE._ns("up_1") == 2 and IE._ns("layer4.0") == 2 and IE._ns("layer1.0") is None and IE._ns("conv1") is None
    finally:
        IE.NS_MAP = saved
    assert [nhwc.pad8(c) for c in (3, 8, 67, 131, 259)] == [8, 8, 72, 136, 264]
    assert nhwc.pick_box(1, 32768) == (128, 1) and nhwc.pick_box(24, 24) == (8, 8) and nhwc.pick_box(48, 48) == (16, 8)
    u = nhwc.ConvUnit(torch.zeros(4, 4, 1, 1), None, None, nhwc.ACT_RELU, nsplit=2, nsplit_out=2)
    assert (u.ns, u.ns_out) == (2, 2) and nhwc.ConvUnit(torch.zeros(4, 4, 1, 1), None, None, nhwc.ACT_RELU).ns == nhwc.NSPLIT
    assert len(IE._sz4([1, 2, 3, 6])) == 4 and [v.value for v in IE._sz4([6])] == [6, 0, 0, 0]


def test_solver_schedules_match_torch_cyclic_lr_and_the_reference_momentum_lambda():
    """istnet_b200.solver evaluates the reference's schedulers in closed form on the host (the values then travel as device scalars):
    CyclicLR(base 1e-5, max 1e-3, triangular, step_size_up = max_epoch * num_mini_batch_per_epoch // 6; utils/solver.py:46-47) and
    the BNMomentumScheduler lambda (utils/solver.py:49)."""
    import warnings

    import torch

    from istnet_b200.solver import bn_momentum_at, cyclic_lr

    for step_up in (1, 2, 7, 20000):
        opt = torch.optim.Adam([torch.nn.Parameter(torch.zeros(1))], lr=0.01)
        sched = torch.optim.lr_scheduler.CyclicLR(opt, base_lr=1e-5, max_lr=1e-3, step_size_up=step_up, mode="triangular", cycle_momentum=False)
        its = list(range(0, 40)) if step_up < 100 else [0, 1, 9999, 19999, 20000, 20001, 39999, 40000, 40001, 119999]
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for it in its:
                sched.step(it)  # the reference passes the iteration explicitly (utils/solver.py:89)
                assert abs(opt.param_groups[0]["lr"] - cyclic_lr(it, 1e-5, 1e-3, step_up)) <= 1e-12, (step_up, it)
    lmbd = lambda it: max(0.9 * 0.5 ** (int(it / 4000)), 0.01)
    for it in (0, 1, 3999, 4000, 8000, 25999, 26000, 28000, 119999):
        assert bn_momentum_at(it, 0.9, 0.5, 4000, 0.01) == lmbd(it)
    assert bn_momentum_at(119999, 0.9, 0.5, 4000, 0.01) == 0.01


def test_solver_batch_merge_and_log_buffer():
    import torch

    from istnet_b200.solver import LogBuffer, merge_into
    from istnet_b200.synth import make_batch

    syn, real = make_batch(3, 32, 16, seed=1), make_batch(2, 32, 16, seed=2)
    keys = ("rgb", "pts", "choose", "category_label", "qo", "rotation_label", "translation_label", "size_label")
    dst = {k: torch.empty((5,) + tuple(syn[k].shape[1:]), dtype=syn[k].dtype) for k in keys}
    assert merge_into(dst, syn, real, keys) == 3
    for k in keys:
        assert torch.equal(dst[k], torch.cat([syn[k], real[k]], dim=0)), k  # utils/solver.py:163-174
    lb = LogBuffer()
    for v in (1.0, 2.0, 3.0, 4.0):
        lb.update({"loss_all": v, "lr": 0.1})
    lb.average(2)
    assert lb._output["loss_all"] == 3.5 and lb.avg["loss_all"] == 2.5
    lb.clear()
    assert not lb.val_history and not lb._output
