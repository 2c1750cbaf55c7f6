"""Probe: conv(cat([a, b])) as conv(a, Wa) followed by cudnn_convolution_add_relu(b, Wb, z=first) -- same FLOPs and the
same bytes as cat + conv, but the extra traffic sits inside compute-bound convolutions instead of a separate pass."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mrfa_b200  # noqa: E402
torch.backends.cudnn.benchmark = True
dev, CL = "cuda", torch.channels_last

def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n

B = 64
with torch.no_grad():
    for (Ca, Cb, Co, R, tag) in ((128, 128, 256, 256, "inp: 128+128 -> 256 @256"), (96, 64, 128, 256, "cf: 96+64 -> 128 @256"),
                                 (128, 128, 256, 128, "inp @128"), (96, 64, 128, 128, "cf @128")):
        a = torch.randn(B, Ca, R, R, device=dev).contiguous(memory_format=CL)
        b = torch.randn(B, Cb, R, R, device=dev).contiguous(memory_format=CL)
        w = (torch.randn(Co, Ca + Cb, 3, 3, device=dev) * 0.02).contiguous(memory_format=CL)
        bias = torch.randn(Co, device=dev)
        wa, wb = w[:, :Ca].contiguous(memory_format=CL), w[:, Ca:].contiguous(memory_format=CL)
        args = ((1, 1), (1, 1), (1, 1), 1)
        f_cat = lambda: torch.cudnn_convolution_relu(torch.ops.mrfa.cat2(a, b), w, bias, *args)
        def f_split():
            z = torch.cudnn_convolution(a, wa, (1, 1), (1, 1), (1, 1), 1, True, False, True)
            return torch.cudnn_convolution_add_relu(b, wb, z, 1.0, bias, *args)
        ref, got = f_cat(), f_split()
        print(json.dumps({"case": tag, "cat2+conv_ms": round(timeit(f_cat), 4), "split_ms": round(timeit(f_split), 4),
                          "max_abs_diff": float((ref - got).abs().max()), "ref_max": float(ref.abs().max())}), flush=True)
        del a, b, ref, got
