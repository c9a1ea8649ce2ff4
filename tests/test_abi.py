"""The C-ABI library loads and exports every symbol include/istnet_b200.h declares (no compute without a GPU)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "istnet_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(istnet_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from istnet_b200 import _C, build

    build.build()
    lib = _C.lib()
    names = _declared()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/istnet_b200.h but not exported"
    assert lib.istnet_version() == 100
    assert lib.istnet_strerror(0) == b"ok"


def test_wrappers_reject_cpu_and_bad_dtypes():
    from istnet_b200 import ext

    with pytest.raises(RuntimeError, match="CPU not supported"):
        ext.furthest_point_sampling(torch.zeros(1, 8, 3), 4)
    with pytest.raises(RuntimeError):
        ext.ball_query(torch.zeros(1, 2, 3), torch.zeros(1, 8, 3), 0.1, 4)


def test_compat_layout_matches_reference_imports():
    import importlib
    import sys

    sys.path.insert(0, os.path.join(ROOT, "compat"))
    try:
        for mod, names in (("ist_net", ("IST_Net", "SupervisedLoss")), ("posenet_gt", ("PoseNetGT", "SupervisedLoss")), ("modules", ("ModifiedResnet", "PointNet2MSG"))):
            m = importlib.import_module(mod)
            for n in names:
                assert hasattr(m, n)
        ext = importlib.import_module("pointnet2._ext")
        for n in ("gather_points", "gather_points_grad", "furthest_point_sampling", "three_nn", "three_interpolate",
                  "three_interpolate_grad", "ball_query", "group_points", "group_points_grad"):
            assert callable(getattr(ext, n))
    finally:
        sys.path.remove(os.path.join(ROOT, "compat"))
        for k in ("ist_net", "posenet_gt", "modules", "pointnet2", "pointnet2._ext"):
            sys.modules.pop(k, None)
