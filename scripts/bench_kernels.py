"""Isolated timing of every hot-path kernel at the bench shapes (B pairs, 256x256 / 512x512).

CUDA events on the launching stream, 3 warm-ups, an L2 flush (256 MB write) between timed
iterations, median of N.  Prints one JSON line per kernel: algorithmic GB/s (or TFLOP/s) and the
fraction of the measured / fallback peak.   python scripts/bench_kernels.py [--batch 64] [--size 256]
"""
import argparse
import json
import os
import statistics
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mrfa_b200                                   # noqa: E402
from mrfa_b200 import _lib, ops                    # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--size", type=int, default=256)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--only", default="")
ap.add_argument("--C", type=int, default=256)
ap.add_argument("--stock", action="store_true", help="also time the stock PyTorch op on the same shapes")
ap.add_argument("--flow", default="smooth", choices=["smooth", "random"],
                help="warp benchmarks: smooth = low-resolution random motion upsampled (what RaftFlow produces); random = i.i.d. per pixel")
a = ap.parse_args()

dev = torch.device("cuda:0")
peaks = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "source": "fallback"}
pp = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(pp):
    d = json.load(open(pp))
    peaks = {"hbm_gbs": float(d.get("hbm_gbs", 6650.0)), "bf16_tflops": float(d.get("bf16_tflops", 1590.0)), "source": "measured"}
flush_buf = torch.empty(256 * 1024 * 1024 // 4, device=dev)


def timeit(fn, iters=a.iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush_buf.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return statistics.median(ts)


def report(name, ms, nbytes=0, flops=0, **extra):
    r = {"kernel": name, "ms": round(ms, 4)}
    if nbytes:
        r["GBps"] = round(nbytes / ms / 1e6, 1)
        r["hbm_frac"] = round(nbytes / ms / 1e6 / peaks["hbm_gbs"], 3)
    if flops:
        r["TFLOPs"] = round(flops / ms / 1e9, 1)
        r["tensor_frac"] = round(flops / ms / 1e9 / peaks["bf16_tflops"], 3)
    r.update(extra)
    print(json.dumps(r), flush=True)


B, S = a.batch, a.size
h = w = S // 4
N, C = h * w, a.C
want = lambda n: (not a.only) or any(t in n or n in t for t in a.only.split(","))

with torch.no_grad():
    if want("membw"):
        big = torch.empty(1 << 30, device=dev, dtype=torch.float32)          # 4 GiB
        ms = timeit(lambda: big.zero_(), 5)
        report("memset_4GiB(write only)", ms, big.numel() * 4)
        half = big[: 1 << 29]
        ms = timeit(lambda: big[1 << 29:].copy_(half), 5)
        report("copy_2GiB(read+write)", ms, big.numel() * 4)
        ms = timeit(lambda: half.sum(), 5)
        report("sum_2GiB(read only)", ms, half.numel() * 4)
        del big, half
    if want("corr"):
        q = torch.randn(B, C, h, w, device=dev)
        k = torch.randn(B, C, h, w, device=dev)
        rows = ops.corr_rows_total(h, w)
        a_op = torch.empty((B, rows, C), device=dev, dtype=torch.bfloat16)
        b_op = torch.empty((B, N, C), device=dev, dtype=torch.bfloat16)
        v0 = torch.empty((B, rows, N), device=dev, dtype=torch.bfloat16)
        v1 = torch.empty((B, rows, N // 4), device=dev, dtype=torch.bfloat16)
        st = lambda: ops._stream()
        layout = ops.corr_map_layout(h, w)
        pack = lambda: ops.check(ops.lib.mrfa_corr_pack(ops._p(q), ops._p(k), ops._p(a_op), ops._p(b_op), B, C, h, w, 0, st()))
        qcl, kcl = q.contiguous(memory_format=torch.channels_last), k.contiguous(memory_format=torch.channels_last)
        pack_cl = lambda: ops.check(ops.lib.mrfa_corr_pack(ops._p(qcl), ops._p(kcl), ops._p(a_op), ops._p(b_op), B, C, h, w, 1, st()))
        ms = timeit(pack_cl)
        report("corr_pack[nhwc]", ms, 8 * q.numel() + 2 * (a_op.numel() + b_op.numel()))
        gemm = lambda: ops.check(ops.lib.mrfa_corr_volume(ops._p(a_op), ops._p(b_op), ops._p(v0), ops._p(v1), B, C, h, w,
                                                          C ** -0.5, ops.sm_count(dev), st()))
        pack()
        ms = timeit(pack)
        report("corr_pack", ms, 8 * q.numel() + 2 * (a_op.numel() + b_op.numel()))
        ms = timeit(gemm)
        report("corr_volume", ms, 2 * (a_op.numel() + b_op.numel() + v0.numel() + v1.numel()), 2 * B * N * N * C,
               epilogue=os.environ.get("MRFA_CORR_EPILOGUE", "tma"), debug=os.environ.get("MRFA_CORR_DEBUG", "0"),
               us_per_pair=round(1e3 * ms / B, 2))
        if a.only == "corr_volume":
            sys.exit(0)
        if a.stock:
            qf, kf = q.flatten(2).transpose(1, 2), k.flatten(2).transpose(1, 2)
            nb = min(B, 16)
            ms = timeit(lambda: torch.einsum("bic,bjc->bij", qf[:nb], kf[:nb]) * C ** -0.5)
            report("stock_einsum_fp32(+scale)", ms * B / nb, 0, 2 * B * N * N * C, note=f"timed on {nb} pairs, scaled")
            # cuBLAS bf16 on the same packed operands: basic-resolution rows only (no pooled rows, no level 1)
            ab, bb = a_op[:, :N], b_op.transpose(1, 2)
            ms = timeit(lambda: torch.bmm(ab, bb))
            report("cublas_bmm_bf16[4096x4096x256 -> bf16]", ms, 2 * (ab.numel() + b_op.numel() + B * N * N), 2 * B * N * N * C,
                   note="plain product: 0.60x of our output bytes, 0.75x of our executed FLOPs, no scale / pooling")
            ms = timeit(lambda: torch.bmm(a_op, bb))
            report("cublas_bmm_bf16[5440x4096x256 -> bf16]", ms, 2 * (a_op.numel() + b_op.numel() + B * rows * N), 2 * B * N * N * C,
                   note="all driving levels, still without the source-pooled level 1 (0.8x of our output bytes)")
            # the library route to the SAME output (all driving levels, scaled, plus the 2x2 source-pooled level 1): the scale
            # folded into the A operand (free), one cuBLAS bf16 GEMM, one avg_pool2d pass over its result
            a_sc = (a_op.float() * C ** -0.5).to(torch.bfloat16)

            def same_output():
                c0 = torch.bmm(a_sc, bb)
                return c0, F.avg_pool2d(c0.view(B * rows, 1, h, w), 2)
            ms_same = timeit(same_output)
            report("cublas_bmm_bf16 + avg_pool2d [same output as corr_volume]", ms_same,
                   2 * (a_op.numel() + b_op.numel() + v0.numel() + v1.numel()) + 2 * v0.numel(), 2 * B * N * N * C,
                   note="bytes include the re-read of level 0 by the pooling pass; row-major maps (ours are 4x8-tiled)")
        # lookups at the six levels
        for i, R in enumerate([S // 32 * 2 ** j for j in range(6)]):
            Rq = min(R, h)
            lvl = max(3 - i, 0)
            coords = (torch.rand(B, 2, Rq, Rq, device=dev) * (h + 4) - 2)
            off = ops.corr_row_offset(h, w, lvl)
            for cl in (False, True):
                fn = lambda: torch.ops.mrfa.corr_lookup(v0, v1, coords, h, w, rows, off, 3, layout, cl)
                ms = timeit(fn)
                report(f"corr_lookup_fwd[Q={Rq}x{Rq} {'nhwc' if cl else 'nchw'}]", ms, B * Rq * Rq * (2 * 64 * 2 + 8 + 98 * 4),
                       layout="tiled" if layout else "rowmajor")
            if a.stock and Rq == h:
                # the reference's CorrBlock on the same GPU: fp32 volume rows -> avg_pool2d level 1 -> two F.grid_sample calls
                # (raft.py:12-48), timed on a few pairs and scaled
                nb = min(B, 4)
                corr32 = v0[:nb, off:off + Rq * Rq].float().reshape(nb * Rq * Rq, 1, h, w)
                cc = coords[:nb]

                def stock_block():
                    lv1 = F.avg_pool2d(corr32, 2, stride=2)
                    d = torch.linspace(-3, 3, 7, device=dev)
                    delta = torch.stack(torch.meshgrid(d, d, indexing="ij"), dim=-1).view(1, 7, 7, 2)
                    cen = cc.permute(0, 2, 3, 1).reshape(-1, 1, 1, 2)
                    outs = []
                    for lv, mp in enumerate((corr32, lv1)):
                        pts = cen / 2 ** lv + delta
                        Hm, Wm = mp.shape[-2:]
                        g = torch.cat([2 * pts[..., :1] / (Wm - 1) - 1, 2 * pts[..., 1:] / (Hm - 1) - 1], -1)
                        outs.append(F.grid_sample(mp, g, align_corners=True).view(nb, Rq, Rq, 49))
                    return torch.cat(outs, -1).permute(0, 3, 1, 2).contiguous()
                ms = timeit(stock_block)
                report(f"stock_CorrBlock[Q={Rq}x{Rq}] (avg_pool2d + 2x grid_sample, fp32 volume)", ms * B / nb,
                       B * Rq * Rq * (2 * 64 * 2 + 8 + 98 * 4), note=f"timed on {nb} pairs, scaled; same algorithmic bytes as ours")
        del v0, v1, a_op, b_op

    if want("warp"):
        chans = (512, 512, 512, 256, 128, 64)
        for i, R in enumerate([S // 32 * 2 ** j for j in range(6)]):
            Cc = chans[i]
            feat = torch.randn(B, Cc, R, R, device=dev)
            if a.flow == "random":                                     # per-pixel random displacements (no tap reuse)
                flow = torch.randn(B, 2, R, R, device=dev) * 2.0
                prior = (mrfa_b200.make_coordinate_grid((R, R), "torch.cuda.FloatTensor")[None] +
                         torch.randn(B, R, R, 2, device=dev) * 0.05).contiguous()
            else:                                                      # motion fields as the path produces them: low-res, upsampled
                lo = max(2, R // 8)
                flow = F.interpolate(torch.randn(B, 2, lo, lo, device=dev) * 3.0, size=(R, R), mode="bilinear", align_corners=True)
                prior = (mrfa_b200.make_coordinate_grid((R, R), "torch.cuda.FloatTensor")[None] +
                         F.interpolate(torch.randn(B, 2, lo, lo, device=dev) * 0.05, size=(R, R), mode="bilinear",
                                       align_corners=True).permute(0, 2, 3, 1)).contiguous()
            elems = B * Cc * R * R
            for fmt in ("nchw", "nhwc"):
                f_ = feat if fmt == "nchw" else feat.contiguous(memory_format=torch.channels_last)
                try:
                    ms = timeit(lambda: mrfa_b200.warp_by_flow(f_, flow))
                    report(f"grid_sample_fwd[{fmt} C={Cc} R={R}]", ms, 4 * (2 * elems + 2 * B * R * R), flow=a.flow)
                    ms = timeit(lambda: torch.ops.mrfa.dual_warp(f_, flow, prior))
                    report(f"dual_warp_fwd[{fmt} C={Cc} R={R}]", ms, 4 * (3 * elems + 4 * B * R * R), flow=a.flow)
                except Exception as e:  # layout not supported yet
                    print(json.dumps({"kernel": f"warp[{fmt} C={Cc} R={R}]", "error": str(e)[:120]}))
            if a.stock:
                ident = mrfa_b200.coords_grid(B, R, R, dev)
                g = (flow + ident).permute(0, 2, 3, 1)
                gn = torch.stack([2 * g[..., 0] / (R - 1) - 1, 2 * g[..., 1] / (R - 1) - 1], -1)
                ms = timeit(lambda: F.grid_sample(feat, gn, align_corners=True))
                report(f"stock_grid_sample[C={Cc} R={R}]", ms, 4 * (2 * elems + 2 * B * R * R))
            del feat

if want("image_warp"):
    # the full-resolution image warp (raft.py:302): NCHW, 3 channels
    img = torch.rand(B, 3, S, S, device=dev)
    flow = F.interpolate(torch.randn(B, 2, S // 8, S // 8, device=dev) * 3.0, size=(S, S), mode="bilinear", align_corners=True)
    ms = timeit(lambda: mrfa_b200.warp_by_flow(img, flow))
    report(f"grid_sample_fwd[image warp, nchw C=3 R={S}]", ms, 4 * (2 * img.numel() + 2 * B * S * S))
    if a.stock:
        g = (flow + mrfa_b200.coords_grid(B, S, S, dev)).permute(0, 2, 3, 1)
        gn = torch.stack([2 * g[..., 0] / (S - 1) - 1, 2 * g[..., 1] / (S - 1) - 1], -1)
        ms = timeit(lambda: F.grid_sample(img, gn, align_corners=True))
        report(f"stock_grid_sample[image warp, C=3 R={S}]", ms, 4 * (2 * img.numel() + 2 * B * S * S))
    del img, flow

if want("bwd"):
    # backward kernels of the training-step configuration (B = 16 per GPU)
    Bt = min(B, 16)
    qb = torch.randn(Bt, C, h, w, device=dev).contiguous(memory_format=torch.channels_last)
    kb = torch.randn(Bt, C, h, w, device=dev).contiguous(memory_format=torch.channels_last)
    rows = ops.corr_rows_total(h, w)
    g0 = torch.randn(Bt, rows, N, device=dev)
    g1 = torch.randn(Bt, rows, N // 4, device=dev)
    with torch.no_grad(), ops.KernelTimer() as kt:
        for _ in range(5):
            torch.ops.mrfa.corr_pyramid_bwd(g0, g1, qb, kb, C ** -0.5)
    for name, r in kt.summary().items():
        ms = r["total_ms"] / r["calls"]
        report(f"{name}[bwd, B={Bt}]", ms, r["bytes"] // r["calls"], r["flops"] // r["calls"], note="in sequence, no L2 flush")
    a_op, b_op = ops.corr_pack_debug(qb, kb)
    gb = (g0 * 0.0625).to(torch.bfloat16)
    with torch.no_grad():
        ms = timeit(lambda: (torch.bmm(gb, b_op), torch.bmm(gb.transpose(1, 2), a_op)))
    report(f"cublas_bmm_bf16 x2 [dA, dB; B={Bt}] (round-1 path, GEMMs only)", ms, 0, 2 * 2 * Bt * rows * N * C)
    del g0, g1, gb
    for (Cc, R) in ((64, 256), (128, 128), (256, 64)):
        feat = torch.randn(Bt, Cc, R, R, device=dev).contiguous(memory_format=torch.channels_last).requires_grad_()
        lo = max(2, R // 8)
        flow = F.interpolate(torch.randn(Bt, 2, lo, lo, device=dev) * 3.0, size=(R, R), mode="bilinear", align_corners=True).requires_grad_()
        go = torch.randn(Bt, Cc, R, R, device=dev).contiguous(memory_format=torch.channels_last)
        grid = flow.permute(0, 2, 3, 1)
        fn = lambda: torch.ops.mrfa.grid_sample_bwd(go, feat.detach(), grid.detach(), _lib.COORD_PIXEL, _lib.PAD_ZEROS, True, 1, True, True)
        ms = timeit(fn)
        elems = Bt * Cc * R * R
        report(f"grid_sample_bwd[nhwc C={Cc} R={R}]", ms, 4 * (4 * elems + 4 * Bt * R * R), run=os.environ.get("MRFA_BWD_RUN", "1"),
               note="bytes: grad_out + input read, grad_input zero-fill + scatter")
        if a.stock:
            ident = mrfa_b200.coords_grid(Bt, R, R, dev)
            gg = (flow.detach() + ident).permute(0, 2, 3, 1)
            gn = torch.stack([2 * gg[..., 0] / (R - 1) - 1, 2 * gg[..., 1] / (R - 1) - 1], -1).requires_grad_()
            f2 = feat.detach().contiguous().requires_grad_()

            def stock():
                o = F.grid_sample(f2, gn, align_corners=True)
                torch.autograd.grad(o, (f2, gn), go.contiguous())
            ms = timeit(stock)
            report(f"stock_grid_sample fwd+bwd[C={Cc} R={R}]", ms, 0)

with torch.no_grad():
    if want("prior"):
        src = torch.rand(B, 3, h, w, device=dev)
        kp_s = torch.rand(B, 10, 2, device=dev) * 1.6 - 0.8
        kp_d = kp_s + torch.randn(B, 10, 2, device=dev) * 0.1
        jac = torch.eye(2, device=dev).view(1, 1, 2, 2) + 0.1 * torch.randn(B, 10, 2, 2, device=dev)
        ms = timeit(lambda: torch.ops.mrfa.dense_motion_prior(kp_d, kp_s, jac, jac, None, src, 0.01))
        report("dense_motion_prior", ms, 4 * B * h * w * (11 * 2 + 11 * 4 + 3))
        pos = torch.randn(1, 10, h, w, device=dev)
        ms = timeit(lambda: torch.ops.mrfa.kp2gaussian(kp_s, pos, h, w, 0.1))
        report("kp2gaussian(+pos)", ms, 4 * B * 10 * h * w)
