"""Probe: the generator's final 7x7 64->3 convolution as a 3x3 convolution over a 4x4 space-to-depth view
(1024 -> 48 channels at 1/4 resolution): same arithmetic, zero taps added, but an N the library kernels can tile."""
import json, os, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
torch.backends.cudnn.benchmark = True
dev = "cuda"
CL = torch.channels_last

def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n

def s2d_weights(w, r=4):
    """(Co, Ci, 7, 7) -> (Co*r*r, r*r*Ci, 3, 3): out channel co*r*r + oy*r + ox, in channel (iy*r + ix)*Ci + ci."""
    Co, Ci, K, _ = w.shape
    w2 = w.new_zeros(Co, r, r, r, r, Ci, 3, 3)
    for by in range(3):
        for iy in range(r):
            for oy in range(r):
                ky = r * (by - 1) + iy - oy + K // 2
                if not 0 <= ky < K: continue
                for bx in range(3):
                    for ix in range(r):
                        for ox in range(r):
                            kx = r * (bx - 1) + ix - ox + K // 2
                            if 0 <= kx < K:
                                w2[:, oy, ox, iy, ix, :, by, bx] = w[:, :, ky, kx]
    return w2.reshape(Co * r * r, r * r * Ci, 3, 3)

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
with torch.no_grad():
    x = torch.randn(B, 64, 256, 256, device=dev).contiguous(memory_format=CL)
    w = (torch.randn(3, 64, 7, 7, device=dev) * 0.02)
    b = torch.randn(3, device=dev)
    ref = F.conv2d(x, w.contiguous(memory_format=CL), b, padding=3)
    print(json.dumps({"case": "final 7x7 64->3 (cuDNN)", "ms": round(timeit(lambda: F.conv2d(x, w.contiguous(memory_format=CL), b, padding=3)), 4)}))
    w2 = s2d_weights(w).contiguous(memory_format=CL)
    b2 = b.repeat_interleave(16)
    s2d = lambda t: t.permute(0, 2, 3, 1).reshape(B, 64, 4, 64, 4, 64).permute(0, 1, 3, 2, 4, 5).reshape(B, 64, 64, 1024).permute(0, 3, 1, 2)
    xs = s2d(x)
    assert xs.is_contiguous(memory_format=CL) or True
    xs = xs.contiguous(memory_format=CL)
    out = F.pixel_shuffle(F.conv2d(xs, w2, b2, padding=1), 4)
    print(json.dumps({"case": "s2d-4 3x3 1024->48 + pixel_shuffle", "max_abs_diff": float((out - ref).abs().max()), "ref_max": float(ref.abs().max()),
                      "conv_ms": round(timeit(lambda: F.conv2d(xs, w2, b2, padding=1)), 4),
                      "s2d_copy_ms": round(timeit(lambda: s2d(x).contiguous(memory_format=CL)), 4),
                      "shuffle_ms": round(timeit(lambda: F.pixel_shuffle(F.conv2d(xs, w2, b2, padding=1), 4)), 4)}))
    w64 = F.pad(w2, (0, 0, 0, 0, 0, 0, 0, 16)).contiguous(memory_format=CL)
    print(json.dumps({"case": "s2d-4 3x3 1024->64 (zero filters)", "conv_ms": round(timeit(lambda: F.conv2d(xs, w64, None, padding=1)), 4)}))
