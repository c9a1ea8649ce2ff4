"""Drop-in for the reference's compiled `pointnet2._ext` (bindings.cpp:11-24): put `compat/` on sys.path
(in place of, or ahead of, the site-packages install of the reference extension) and the reference's
`model/pointnet2/pointnet2_utils.py` imports these B200 kernels unchanged."""
from istnet_b200.ext import (  # noqa: F401
    ball_query,
    furthest_point_sampling,
    gather_points,
    gather_points_grad,
    group_points,
    group_points_grad,
    three_interpolate,
    three_interpolate_grad,
    three_nn,
)
