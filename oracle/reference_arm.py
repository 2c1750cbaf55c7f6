"""Run the UNMODIFIED reference (vendored by oracle/build_ref.py into oracle/_ref/) on the CPU.

TEST INFRASTRUCTURE ONLY: used by tests/, by bench.py's `cpu_baseline` and `--impl reference`
legs, never by the product.  This module imports neither ``mrfa_b200`` nor anything that loads
libmrfa_b200.so; it composes the path exactly as the reference's own callers do
(demo.py:47-73, model.py:201,210):

    dense_motion = DenseMotionNetwork(source, kp_driving, kp_source, bg_param)      # dense_motion.py:104
    out, warp_img, occlusion = RaftFlow(kp_s, kp_d, dense_motion,
                                        img=AntiAliasInterpolation2d(3, 0.25)(source), img_full=source)   # raft.py:141

The key-point detector is upstream of the hot path and is replaced by
synthetic key-points in every arm of the benchmark.
"""
from __future__ import annotations

import importlib
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


def available() -> bool:
    return os.path.exists(os.path.join(REF_DIR, "modules", "raft.py"))


_MODS = None


def load():
    """Import the vendored reference modules (util, raft, dense_motion) from oracle/_ref."""
    global _MODS
    if _MODS is not None:
        return _MODS
    if not available():
        raise ImportError("oracle/_ref is missing: run `python oracle/build_ref.py` in the build container "
                          "(it copies /root/reference/{modules,config}; the copy ships to the GPU box)")
    for p in (os.path.join(REF_DIR, "_shims"), REF_DIR):
        if p not in sys.path:
            sys.path.insert(0, p)
    mods = {n: importlib.import_module("modules." + n) for n in ("util", "raft", "dense_motion")}
    for m in mods.values():                                 # the modules that ran are the vendored ones
        assert os.path.abspath(m.__file__).startswith(REF_DIR), m.__file__
    _MODS = mods
    return mods


def build_networks(cfg: dict, size: int, prior: str = "fomm"):
    """(down, dense_motion, raft_flow) of the reference, constructed from the reference YAML dict."""
    m = load()
    if prior == "tpsm":
        dm = m["dense_motion"].TPSDenseMotionNetwork(**cfg["tpsm_dense_motion"])
    else:
        dm = m["dense_motion"].DenseMotionNetwork(**cfg["dense_motion"])
    rf = m["raft"].RaftFlow(**dict(cfg["raft_flow"], size=size))
    down = m["util"].AntiAliasInterpolation2d(3, 0.25)       # model.py:170 / demo.py:25
    return down.eval(), dm.eval(), rf.eval()


def forward(nets, src, kp_s, kp_d, bg=None):
    """One refinement forward; returns (out, warp_img, occlusion, dense_motion dict)."""
    down, dm, rf = nets
    dense = dm(src, kp_d, kp_s, bg_param=bg)
    out, warp_img, occ = rf(kp_s["kp"], kp_d["kp"], dense, img=down(src), img_full=src)
    return out, warp_img, occ, dense
