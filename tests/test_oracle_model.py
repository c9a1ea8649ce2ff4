"""The oracle's torch port (oracle/istnet_port.py) against (a) the golden vectors produced by the reference itself
and (b) — where /root/reference exists — the unmodified reference modules run live."""
import numpy as np
import pytest
import torch

from conftest import fixed_dropout_noise, golden_inputs, load_golden, rel_err, sd_checksum
from istnet_b200 import model as M
from oracle import istnet_port as port
from oracle import ref_harness

LABELS = ("qo", "rotation_label", "translation_label", "size_label")


def _sd(model, grad=False):
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    if grad:
        for k, v in sd.items():
            if v.is_floating_point() and "running_" not in k:
                v.requires_grad_(True)
    return sd


def test_port_reproduces_cfg0_golden():
    z = load_golden("cfg0_eval.npz")
    torch.manual_seed(1)
    m = M.IST_Net(6, False)
    assert abs(sd_checksum(m.state_dict()) - float(z["sd_checksum"])) < 1e-6 * float(z["sd_checksum"])
    with torch.no_grad():
        ep = port.ist_net_forward(_sd(m), golden_inputs(z), training=False)
    assert set(ep) == {"pred_qo", "pred_rotation", "pred_translation", "pred_size"}
    for k, v in ep.items():
        assert rel_err(v, z["out_" + k]) < 1e-6, k


def test_port_reproduces_cfg4_dense_cloud_golden():
    """4096-point crop (BASELINE.json configs[4] shape), eval forward with non-trivial BatchNorm statistics: port == reference."""
    from conftest import perturb_batchnorm

    z = load_golden("cfg4_eval.npz")
    torch.manual_seed(1)
    m = M.IST_Net(6, False)
    perturb_batchnorm(m, seed=41)
    assert abs(sd_checksum(m.state_dict()) - float(z["sd_checksum"])) < 1e-6 * float(z["sd_checksum"])
    with torch.no_grad():
        ep = port.ist_net_forward(_sd(m), golden_inputs(z), training=False)
    for k, v in ep.items():
        assert rel_err(v, z["out_" + k]) < 1e-6, k


def test_port_reproduces_train_golden():
    z = load_golden("train_b4.npz")
    torch.manual_seed(1)
    m = M.IST_Net(6, False)
    sd = _sd(m, grad=True)
    inp = golden_inputs(z)
    noise = fixed_dropout_noise(77)
    b = inp["rgb"].shape[0]
    masks = [noise(b, 1024, 0.3), noise(b, 256, 0.15), noise(b, 64, 0.15)]
    ep = port.ist_net_forward(sd, inp, training=True, bn_momentum=0.9, dropout_noise=masks)
    for k, v in ep.items():
        assert rel_err(v, z["out_" + k]) < 1e-6, k
    loss = port.ist_net_loss(ep, inp)
    assert abs(loss.item() - float(z["loss"])) < 1e-6 * abs(float(z["loss"]))
    loss.backward()
    for k in z:
        if k.startswith("gradnorm_"):
            g = sd[k[9:]].grad
            assert abs(g.double().norm().item() - float(z[k])) <= 1e-5 * float(z[k]) + 1e-12, k
        elif k.startswith("stat_"):
            assert rel_err(sd[k[5:]], z[k]) < 1e-6, k
    for n in z["nograd"]:
        assert sd[str(n)].grad is None


@pytest.mark.skipif(not ref_harness.available(), reason="reference tree not present on this host")
def test_port_equals_live_reference_posenet_gt():
    ns = ref_harness.load()
    torch.manual_seed(1)
    mine = M.PoseNetGT(6)
    ref = ns.posenet_gt.PoseNetGT(6)
    ref.load_state_dict(mine.state_dict())
    ref.eval()
    from istnet_b200.synth import make_batch

    inp = make_batch(1, 256, 64, seed=5, quantize=True)
    with torch.no_grad():
        a = ref({k: v.clone() for k, v in inp.items()})
        b = port.posenet_gt_forward(_sd(mine), inp, training=False)
    for k in a:
        assert torch.equal(a[k], b[k]), k


def test_port_reproduces_posenet_gt_golden():
    z = load_golden("posenet_gt_b2.npz")
    torch.manual_seed(1)
    m = M.PoseNetGT(6)
    sd = _sd(m, grad=True)
    inp = golden_inputs(z)
    noise = fixed_dropout_noise(78)
    masks = [noise(2, 1024, 0.3), noise(2, 256, 0.15), noise(2, 64, 0.15)]
    ep = port.posenet_gt_forward(sd, inp, training=True, dropout_noise=masks)
    for k, v in ep.items():
        assert rel_err(v, z["out_" + k]) < 1e-6, k
    loss = port.posenet_gt_loss(ep, inp)
    loss.backward()
    for n in z["nograd"]:
        assert sd[str(n)].grad is None, n
    for k in z:
        if k.startswith("gradnorm_"):
            assert abs(sd[k[9:]].grad.double().norm().item() - float(z[k])) <= 1e-5 * float(z[k]) + 1e-12, k
