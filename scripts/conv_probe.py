"""A/B probe for the small-channel convolutions around the hot path (7x7 convs with 2/3 input or 3 output
channels).  cuDNN serves them with a legacy indexed kernel or a 64-wide tile for 3 outputs; this times the
zero-padded-channel variants and the GEMM + shift-add decomposition of the final 7x7 convolution."""
import json
import sys
import os

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
torch.backends.cudnn.benchmark = True
dev = "cuda"
CL = torch.channels_last


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n


def rep(name, ms, **kw):
    print(json.dumps({"case": name, "ms": round(ms, 4), **kw}), flush=True)


B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
with torch.no_grad():
    for (cin, cout, R, tag) in ((2, 128, 256, "convf1@256"), (2, 128, 128, "convf1@128"), (3, 64, 256, "first@256")):
        x = torch.randn(B, cin, R, R, device=dev).contiguous(memory_format=CL)
        w = (torch.randn(cout, cin, 7, 7, device=dev) * 0.05).contiguous(memory_format=CL)
        b = torch.randn(cout, device=dev)
        ref = torch.cudnn_convolution_relu(x, w, b, (1, 1), (3, 3), (1, 1), 1)
        rep(tag + " cin=%d" % cin, timeit(lambda: torch.cudnn_convolution_relu(x, w, b, (1, 1), (3, 3), (1, 1), 1)))
        import mrfa_b200
        wpk = mrfa_b200.ops.conv7x7_small_pack(w)
        out = torch.ops.mrfa.conv7x7_small(x, wpk, b, True)
        ms = timeit(lambda: torch.ops.mrfa.conv7x7_small(x, wpk, b, True))
        rep(tag + " mrfa::conv7x7_small (tcgen05 tf32)", ms, max_abs_diff=float((out - ref).abs().max()),
            out_GBps=round(out.numel() * 4 / ms / 1e6, 1))
        if os.environ.get("PROBE_PAD", "0") == "0":
            del x, ref
            continue
        for pc in (4, 8):
            xp = F.pad(x, (0, 0, 0, 0, 0, pc - cin)).contiguous(memory_format=CL)
            wp = F.pad(w, (0, 0, 0, 0, 0, pc - cin)).contiguous(memory_format=CL)
            out = torch.cudnn_convolution_relu(xp, wp, b, (1, 1), (3, 3), (1, 1), 1)
            err = float((out - ref).abs().max())
            rep(tag + " padded cin=%d" % pc, timeit(lambda: torch.cudnn_convolution_relu(xp, wp, b, (1, 1), (3, 3), (1, 1), 1)),
                max_abs_diff=err, pad_ms=round(timeit(lambda: F.pad(x, (0, 0, 0, 0, 0, pc - cin)).contiguous(memory_format=CL)), 4))
        del x, ref
    if os.environ.get("PROBE_FINAL", "0") == "0":
        sys.exit(0)
    # final 7x7 64 -> 3
    x = torch.randn(B, 64, 256, 256, device=dev).contiguous(memory_format=CL)
    w = (torch.randn(3, 64, 7, 7, device=dev) * 0.02).contiguous(memory_format=CL)
    b = torch.randn(3, device=dev)
    ref = F.conv2d(x, w, b, padding=3)
    rep("final 64->3", timeit(lambda: F.conv2d(x, w, b, padding=3)))
    for pc in (4, 8, 16):
        wp = F.pad(w, (0, 0, 0, 0, 0, 0, 0, pc - 3)).contiguous(memory_format=CL)
        bp = F.pad(b, (0, pc - 3))
        out = F.conv2d(x, wp, bp, padding=3)[:, :3]
        rep("final 64->%d (zero filters)" % pc, timeit(lambda: F.conv2d(x, wp, bp, padding=3)), max_abs_diff=float((out - ref).abs().max()))
    # GEMM half of the decomposition: Yt[b, tap*3+co, p] = sum_ci W[co, ci, tap] * x[b, ci, p]
    wt = w.permute(2, 3, 0, 1).reshape(147, 64).contiguous()
    xv = x.permute(0, 2, 3, 1).reshape(B, 256 * 256, 64)              # NHWC memory viewed as (B, px, 64)
    rep("final: GEMM (147x64)x(64xpx) -> planes", timeit(lambda: torch.matmul(wt, xv.transpose(1, 2))))
    rep("final: GEMM (px x64)x(64x147) -> NHWC", timeit(lambda: torch.matmul(xv, wt.t())))
    wt160 = F.pad(wt, (0, 0, 0, 13))
    rep("final: GEMM (160x64)x(64xpx) -> planes", timeit(lambda: torch.matmul(wt160, xv.transpose(1, 2))))
    y = torch.matmul(wt, xv.transpose(1, 2))
    rep("stream read of the planes (sum)", timeit(lambda: y.sum()))
