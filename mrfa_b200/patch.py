"""``patch_reference()``: rebind the hot-path names inside an imported MRFA checkout.

The reference modules use ``from .util import ...`` so patching ``modules.util`` alone is not
enough; the names are rebound inside ``modules.raft`` and ``modules.dense_motion`` as well, and
the network classes themselves are swapped because their bare ``F.grid_sample`` calls
(raft.py:166,168,271; dense_motion.py:83,241) cannot be intercepted by name.
"""
from __future__ import annotations

import importlib


def patch_reference(modules_pkg: str = "modules"):
    from . import corr, equivariance, prior_motion, refine, sampling

    util = importlib.import_module(modules_pkg + ".util")
    raft = importlib.import_module(modules_pkg + ".raft")
    dense = importlib.import_module(modules_pkg + ".dense_motion")
    fns = ("bilinear_sampler", "batch_bilinear_sampler", "coords_grid", "kp2gaussian", "make_coordinate_grid",
           "to_homogeneous", "from_homogeneous", "TPS")
    for mod in (util, raft, dense):
        for name in fns:
            if hasattr(mod, name):
                setattr(mod, name, getattr(sampling, name))
    util.deform_input = sampling.deform_input
    raft.CorrBlock = corr.CorrBlock
    raft.RaftFlow = refine.RaftFlow
    raft.BasicMotionEncoder = refine.BasicMotionEncoder
    raft.RefineFlow = refine.RefineFlow
    dense.DenseMotionNetwork = prior_motion.DenseMotionNetwork
    dense.TPSDenseMotionNetwork = prior_motion.TPSDenseMotionNetwork
    try:                                   # modules.model imports the classes by name at import time
        model = importlib.import_module(modules_pkg + ".model")
        model.RaftFlow = refine.RaftFlow
        model.DenseMotionNetwork = prior_motion.DenseMotionNetwork
        model.TPSDenseMotionNetwork = prior_motion.TPSDenseMotionNetwork
        model.Transform = equivariance.Transform
    except Exception:                      # model.py needs torchvision weights / .cuda(); optional
        pass
    return {"util": util, "raft": raft, "dense_motion": dense}
