"""Flat-import shim for `modules` (reference model/modules.py): ModifiedResnet, PointNet2MSG."""
from istnet_b200.image import Modified_PSPNet, ModifiedResnet, PSPModule, PSPUpsample  # noqa: F401
from istnet_b200.pointnet2 import PointNet2MSG  # noqa: F401
