"""Flat-import shim: `from ist_net import IST_Net, SupervisedLoss` (reference train.py:78)."""
from istnet_b200.model import IST_Net, SupervisedLoss, FeatureDeformer, HeavyEstimator, ImplicitTransformation, LightEstimator, WorldSpaceEnhancer  # noqa: F401
