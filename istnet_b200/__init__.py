"""istnet_b200 — B200-native (sm_100a) implementation of IST-Net's per-instance forward/backward hot path.

Layout (tier: ONE hot path):
  csrc/          hand-written CUDA kernels + the C ABI (include/istnet_b200.h) -> libistnet_b200.so
  _C.py, ext.py  ctypes binding; the nine `pointnet2._ext` operators (reference bindings.cpp:11-24)
  functional.py  autograd front-ends of the kernels
  pointnet2.py, image.py, model.py   the reference's module surface (IST_Net, PoseNetGT, ...), same state_dict keys
  parallel.py    one-process-per-GPU gradient all-reduce (NCCL)
  synth.py       synthetic RGB-D instance crops (bench / tests)
"""
# Importing this package changes NO global PyTorch state.  The 1e-4 float parity target (BASELINE.json north_star) is met by the
# package's own kernels; the few small library matmuls left on the path (PSP prior maps, `nn.Linear` pose heads) need
# `torch.backends.cuda.matmul.allow_tf32 == False`, which is PyTorch's default — `model.check_fp32_matmul()` raises if a host
# program turned TF32 matmuls on.
from .model import IST_Net, PoseNetGT, SupervisedLoss, PoseNetGTLoss, LossCfg  # noqa: E402,F401
