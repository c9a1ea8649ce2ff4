import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def rel_err(a, b):
    """max|a-b| / max|b| — the float parity measure (north_star: <= 1e-4)."""
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name), allow_pickle=False)
    return {k: z[k] for k in z.files}


def golden_inputs(z):
    return {k[3:]: torch.from_numpy(v) for k, v in z.items() if k.startswith("in_")}


def sd_checksum(sd):
    return float(sum(v.double().abs().sum().item() for v in sd.values() if v.is_floating_point()))


def fixed_dropout_noise(seed):
    g = torch.Generator().manual_seed(seed)

    def fn(b, c, p):
        return torch.empty(b, c, 1, 1).bernoulli_(1 - p, generator=g).div_(1 - p)

    return fn


def perturb_batchnorm(model, seed):
    """Deterministic non-trivial BatchNorm running statistics / affine parameters (eval-mode goldens would otherwise see
    mean 0, var 1, gamma 1, beta 0 everywhere).  Used by the golden generator and by the tests on the same module tree."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for mod in model.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.running_mean.copy_(0.1 * torch.randn(mod.num_features, generator=g))
                mod.running_var.copy_(0.5 + torch.rand(mod.num_features, generator=g))
                mod.weight.copy_(0.5 + torch.rand(mod.num_features, generator=g))
                mod.bias.copy_(0.1 * torch.randn(mod.num_features, generator=g))
