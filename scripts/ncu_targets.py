"""Launch each hot-path kernel a few times at the bench shapes so `ncu -k regex:<name>` can capture it in isolation.

    ncu --set full --clock-control none --import-source on -k regex:corr_lookup_fwd_tiled -s 2 -c 1 -o gpurun_out/x python scripts/ncu_targets.py --only lookup
"""
import argparse
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mrfa_b200                                   # noqa: E402
from mrfa_b200 import ops                          # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--only", default="lookup,warp,corr,bwd")
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
dev = torch.device("cuda:0")
B, h, w, C = a.batch, 64, 64, 256
want = lambda n: n in a.only.split(",")
torch.manual_seed(0)
with torch.no_grad():
    if want("corr") or want("lookup"):
        q = torch.randn(B, C, h, w, device=dev).contiguous(memory_format=torch.channels_last)
        k = torch.randn(B, C, h, w, device=dev).contiguous(memory_format=torch.channels_last)
        for _ in range(a.reps):
            pyr = mrfa_b200.CorrPyramid(q, k, C ** -0.5)
        if want("lookup"):
            # coordinates as the path produces them: identity + a smooth flow of a few pixels
            flow = F.interpolate(torch.randn(B, 2, 8, 8, device=dev) * 3.0, size=(h, w), mode="bilinear", align_corners=True)
            coords = flow + mrfa_b200.coords_grid(B, h, w, dev)
            for _ in range(a.reps):
                pyr.block(0)(coords, True)
    if want("warp"):
        R, Cc = 256, 64
        feat = torch.randn(B, Cc, R, R, device=dev).contiguous(memory_format=torch.channels_last)
        flow = F.interpolate(torch.randn(B, 2, 32, 32, device=dev) * 3.0, size=(R, R), mode="bilinear", align_corners=True)
        prior = (mrfa_b200.make_coordinate_grid((R, R), "torch.cuda.FloatTensor")[None] +
                 F.interpolate(torch.randn(B, 2, 32, 32, device=dev) * 0.05, size=(R, R), mode="bilinear",
                               align_corners=True).permute(0, 2, 3, 1)).contiguous()
        for _ in range(a.reps):
            torch.ops.mrfa.dual_warp(feat, flow, prior)
            mrfa_b200.warp_by_flow(feat, flow)
    if want("image"):
        # the full-resolution image warp (raft.py:302): NCHW, 3 channels -> grid_sample_fwd_fewc_kernel; stock op beside it
        img = torch.rand(B, 3, 256, 256, device=dev)
        flow = F.interpolate(torch.randn(B, 2, 32, 32, device=dev) * 3.0, size=(256, 256), mode="bilinear", align_corners=True)
        g = (flow + mrfa_b200.coords_grid(B, 256, 256, dev)).permute(0, 2, 3, 1)
        gn = torch.stack([2 * g[..., 0] / 255 - 1, 2 * g[..., 1] / 255 - 1], -1)
        for _ in range(a.reps):
            mrfa_b200.warp_by_flow(img, flow)
            F.grid_sample(img, gn, align_corners=True)
    if want("prior"):
        import synthetic_inputs as syn
        kp_s, kp_d = ({k: v.to(dev) for k, v in d.items()} for d in syn.keypoints(B, 10, seed=0))
        src = torch.rand(B, 3, 64, 64, device=dev)
        for _ in range(a.reps):
            torch.ops.mrfa.dense_motion_prior(kp_d["kp"], kp_s["kp"], kp_d["jacobian"], kp_s["jacobian"], None, src, 0.01)
if want("bwd"):
    Bt = 16
    q = torch.randn(Bt, C, h, w, device=dev).contiguous(memory_format=torch.channels_last)
    k = torch.randn(Bt, C, h, w, device=dev).contiguous(memory_format=torch.channels_last)
    rows = ops.corr_rows_total(h, w)
    g0 = torch.randn(Bt, rows, h * w, device=dev)
    g1 = torch.randn(Bt, rows, h * w // 4, device=dev)
    for _ in range(a.reps):
        torch.ops.mrfa.corr_pyramid_bwd(g0, g1, q, k, C ** -0.5)
    feat = torch.randn(Bt, 64, 256, 256, device=dev).contiguous(memory_format=torch.channels_last).requires_grad_()
    flow = F.interpolate(torch.randn(Bt, 2, 32, 32, device=dev) * 3.0, size=(256, 256), mode="bilinear", align_corners=True).requires_grad_()
    for _ in range(a.reps):
        out = mrfa_b200.warp_by_flow(feat, flow)
        out.backward(torch.ones_like(out))
    # lookup backward: volume gradients accumulated into the pyramid's fp32 buffers + coordinate gradients
    pyr = mrfa_b200.CorrPyramid(q.detach().requires_grad_(), k.detach().requires_grad_(), C ** -0.5)
    cflow = F.interpolate(torch.randn(Bt, 2, 8, 8, device=dev) * 3.0, size=(h, w), mode="bilinear", align_corners=True)
    coords = (cflow + mrfa_b200.coords_grid(Bt, h, w, dev)).requires_grad_()
    for _ in range(a.reps):
        o = pyr.block(0)(coords, True)
        o.backward(torch.ones_like(o), retain_graph=True)
    # prior-motion synthesis backward (key-points, Jacobians and source carry gradients)
    import synthetic_inputs as syn
    kp_s, kp_d = ({kk: v.to(dev).requires_grad_() for kk, v in d.items()} for d in syn.keypoints(Bt, 10, seed=0))
    src = torch.rand(Bt, 3, 64, 64, device=dev, requires_grad=True)
    for _ in range(a.reps):
        mo, hg = torch.ops.mrfa.dense_motion_prior(kp_d["kp"], kp_s["kp"], kp_d["jacobian"], kp_s["jacobian"], None, src, 0.01)
        (mo.sum() + hg.sum()).backward()
torch.cuda.synchronize()
print("done")
