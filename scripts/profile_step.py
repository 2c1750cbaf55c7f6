"""One refinement-forward step under cudaProfilerStart/Stop (for `ncu --profile-from-start off`).

    ncu --profile-from-start off ... python scripts/profile_step.py [--batch 64] [--size 256]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import mrfa_b200                                   # noqa: E402
import synthetic_inputs as syn            # noqa: E402
import yaml                                         # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--size", type=int, default=256)
ap.add_argument("--warmup", type=int, default=2)
ap.add_argument("--nchw", action="store_true")
a = ap.parse_args()

cfg = yaml.safe_load(open(os.path.join(ROOT, "tests", "golden", "vox1.yaml")))
dev = torch.device("cuda:0")
torch.manual_seed(0)
dm = mrfa_b200.DenseMotionNetwork(**cfg["dense_motion"]).to(dev).eval()
rf = mrfa_b200.RaftFlow(**dict(cfg["raft_flow"], size=a.size)).to(dev).eval()
if not a.nchw:
    dm.channels_last_()
    rf.channels_last_()
src, _ = syn.frame_pairs(a.batch, a.size)
kp_s, kp_d = syn.keypoints(a.batch, 10)
src = src.to(dev)
kp_s = {k: v.to(dev) for k, v in kp_s.items()}
kp_d = {k: v.to(dev) for k, v in kp_d.items()}


def step():
    dense = dm(src, kp_d, kp_s)
    return rf(kp_s["kp"], kp_d["kp"], dense, img=dm.down(src), img_full=src)[0]


with torch.no_grad():
    for _ in range(a.warmup):
        step()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    step()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
print("done")
