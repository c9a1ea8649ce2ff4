"""Flat-import shim: `from posenet_gt import PoseNetGT, SupervisedLoss` (reference train.py:81)."""
from istnet_b200.model import HeavyEstimator, PoseNetGT  # noqa: F401
from istnet_b200.model import PoseNetGTLoss as SupervisedLoss  # noqa: F401
