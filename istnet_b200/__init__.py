"""istnet_b200 — B200-native (sm_100a) implementation of IST-Net's per-instance forward/backward hot path.

Layout (tier: ONE hot path):
  csrc/          hand-written CUDA kernels + the C ABI (include/istnet_b200.h) -> libistnet_b200.so
  _C.py, ext.py  ctypes binding; the nine `pointnet2._ext` operators (reference bindings.cpp:11-24)
  functional.py  autograd front-ends of the kernels
  pointnet2.py, image.py, model.py   the reference's module surface (IST_Net, PoseNetGT, ...), same state_dict keys
  parallel.py    one-process-per-GPU gradient all-reduce (NCCL)
  synth.py       synthetic RGB-D instance crops (bench / tests)
"""
import torch

# Float parity target is 1e-4 relative (BASELINE.json north_star): single-pass TF32 does not meet it, so library
# convolutions / matmuls used by this package run in true FP32.
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False

from .model import IST_Net, PoseNetGT, SupervisedLoss, PoseNetGTLoss, LossCfg  # noqa: E402,F401
