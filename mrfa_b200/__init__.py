"""mrfa_b200 -- B200-native (sm_100a) drop-in for MRFA's per-frame-pair motion-refinement hot
path: CorrBlock + iterative flow-update lookup, kp2gaussian / make_coordinate_grid / TPS
sparse-motion construction, deformed-source stacking and the grid_sample feature warps.

Importing this package loads libmrfa_b200.so and registers the ``torch.ops.mrfa.*`` custom
ops; it raises if the library has not been built (``python -m mrfa_b200.build``).
"""
from . import ops  # noqa: F401  (registers torch.ops.mrfa.*)
from .corr import CorrBlock, CorrPyramid
from .prior_motion import DenseMotionNetwork, TPSDenseMotionNetwork
from .refine import BasicMotionEncoder, RaftFlow, RefineFlow
from .sampling import (TPS, batch_bilinear_sampler, bilinear_sampler, coords_grid, deform_input, from_homogeneous,
                       grid_sample, kp2gaussian, make_coordinate_grid, to_homogeneous, warp_by_flow)
from .blocks import AntiAliasInterpolation2d, Hourglass, OcclusionAwareGenerator
from .equivariance import Transform
from .graphs import GraphedRefiner
from .patch import patch_reference

__all__ = [
    "CorrBlock", "CorrPyramid", "DenseMotionNetwork", "TPSDenseMotionNetwork", "RaftFlow", "BasicMotionEncoder",
    "RefineFlow", "TPS", "batch_bilinear_sampler", "bilinear_sampler", "coords_grid", "deform_input",
    "from_homogeneous", "to_homogeneous", "grid_sample", "kp2gaussian", "make_coordinate_grid", "warp_by_flow",
    "AntiAliasInterpolation2d", "Hourglass", "OcclusionAwareGenerator", "patch_reference", "GraphedRefiner", "Transform",
]
