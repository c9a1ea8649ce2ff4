"""ORACLE — test infrastructure only (see oracle/__init__.py).

Imports the UNMODIFIED reference Python (`/root/reference/model`, `utils`) in-process with the four shims
of SURVEY.md §8(c).  Only usable where /root/reference exists (this container, never the GPU box): it is
used to (a) validate oracle/istnet_port.py and oracle/pointops_ref.c, (b) generate tests/golden/*.

Shims (nothing in /root/reference is modified):
  1. sys.modules['pointnet2._ext'] = oracle.pointops   (reference ops are CUDA-only, sampling.cpp:39)
  2. torch.Tensor.cuda -> identity on a CPU-only host  (ist_net.py:38, rotation_utils.py:6 hard-code .cuda())
  3. resnet.resnet18(pretrained) -> pretrained=False    (modules.py:52-54 would download weights)
  4. sys.path as train.py:11-14
"""
import importlib
import os
import sys

import torch

REF = os.environ.get("ISTNET_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF, "model"))


_loaded = None


def load(ext_module=None):
    """Returns a namespace with the reference modules: ist_net, posenet_gt, modules, losses, pointnet2_utils."""
    global _loaded
    if _loaded is not None:
        return _loaded
    assert available(), "reference tree not present"
    from . import pointops

    for p in ("model", "model/pointnet2", "utils", "provider"):
        sys.path.insert(0, os.path.join(REF, p))
    sys.modules["pointnet2._ext"] = ext_module or pointops
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
    import resnet

    _orig18 = resnet.resnet18
    resnet.resnet18 = lambda pretrained=False: _orig18(False)

    class NS:
        pass

    ns = NS()
    for name in ("ist_net", "posenet_gt", "modules", "losses", "rotation_utils", "pointnet2_modules"):
        setattr(ns, name, importlib.import_module(name))
    ns.pointnet2_utils = importlib.import_module("pointnet2.pointnet2_utils")
    _loaded = ns
    return ns
