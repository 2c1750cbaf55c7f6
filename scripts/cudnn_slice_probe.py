"""Probe: can cuDNN (through its graph front-end) write conv + bias + ReLU straight into a channel slice of a wider
NHWC buffer?  If so the update block's cat([cor, flo]) / cat([motion, context]) passes disappear."""
import json
import os
import sys

import cudnn
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
torch.backends.cudnn.benchmark = True
dev = "cuda"
CL = torch.channels_last
FL = cudnn.data_type.FLOAT


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


def nhwc_strides(N, C, H, W, Ct=None):
    Ct = Ct or C
    return [H * W * Ct, 1, W * Ct, Ct]


def build(handle, N, Cin, Cout, H, W, out_Ct, k=3):
    g = cudnn.pygraph(handle=handle, io_data_type=FL, intermediate_data_type=FL, compute_data_type=FL)
    X = g.tensor(name="X", dim=[N, Cin, H, W], stride=nhwc_strides(N, Cin, H, W), data_type=FL)
    Wt = g.tensor(name="W", dim=[Cout, Cin, k, k], stride=[Cin * k * k, 1, k * Cin, Cin], data_type=FL)
    Bt = g.tensor(name="B", dim=[1, Cout, 1, 1], stride=[Cout, 1, Cout, Cout], data_type=FL)
    c = g.conv_fprop(image=X, weight=Wt, padding=[k // 2, k // 2], stride=[1, 1], dilation=[1, 1])
    bsum = g.bias(name="bias", input=c, bias=Bt)
    Y = g.relu(name="relu", input=bsum)
    Y.set_output(True).set_data_type(FL).set_dim([N, Cout, H, W]).set_stride(nhwc_strides(N, Cout, H, W, out_Ct))
    g.validate()
    g.build_operation_graph()
    g.create_execution_plans([cudnn.heur_mode.A, cudnn.heur_mode.FALLBACK])
    g.check_support()
    g.build_plans(cudnn.build_plan_policy.HEURISTICS_CHOICE)
    return g, X, Wt, Bt, Y


B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
R = int(sys.argv[2]) if len(sys.argv) > 2 else 256
handle = cudnn.create_handle()
cudnn.set_stream(handle=handle, stream=torch.cuda.current_stream().cuda_stream)
with torch.no_grad():
    x = torch.randn(B, 128, R, R, device=dev).contiguous(memory_format=CL)
    buf = torch.zeros(B, 160, R, R, device=dev).contiguous(memory_format=CL)
    for (Cout, off, tag) in ((96, 0, "convc2 128->96 -> buf[:, 0:96]"), (64, 96, "convf2 128->64 -> buf[:, 96:160]")):
        w = (torch.randn(Cout, 128, 3, 3, device=dev) * 0.03).contiguous(memory_format=CL)
        b = torch.randn(Cout, device=dev)
        ref = torch.cudnn_convolution_relu(x, w, b, (1, 1), (1, 1), (1, 1), 1)
        t_ref = timeit(lambda: torch.cudnn_convolution_relu(x, w, b, (1, 1), (1, 1), (1, 1), 1))
        for Ct, name in ((Cout, "dense"), (160, "slice")):
            try:
                g, X, Wt, Bt, Y = build(handle, B, 128, Cout, R, R, Ct)
                ws = torch.empty(max(1, g.get_workspace_size()), device=dev, dtype=torch.uint8)
                dst = torch.empty_like(ref) if Ct == Cout else buf
                ptr = dst.data_ptr() + (0 if Ct == Cout else off * 4)
                run = lambda: g.execute({X: x, Wt: w, Bt: b, Y: ptr}, ws, handle=handle)
                run()
                torch.cuda.synchronize()
                got = dst if Ct == Cout else buf[:, off:off + Cout]
                print(json.dumps({"case": tag, "output": name, "max_abs_diff": float((got - ref).abs().max()), "ref_max": float(ref.abs().max()),
                                  "ms": round(timeit(run), 4), "torch_cudnn_relu_ms": round(t_ref, 4),
                                  "workspace_MB": round(g.get_workspace_size() / 1e6, 1)}), flush=True)
            except Exception as e:  # noqa: BLE001
                print(json.dumps({"case": tag, "output": name, "error": repr(e)[:300]}), flush=True)
    import mrfa_b200
    a96 = torch.randn(B, 96, R, R, device=dev).contiguous(memory_format=CL)
    a64 = torch.randn(B, 64, R, R, device=dev).contiguous(memory_format=CL)
    print(json.dumps({"case": "mrfa::cat2 96+64", "ms": round(timeit(lambda: torch.ops.mrfa.cat2(a96, a64)), 4)}))
