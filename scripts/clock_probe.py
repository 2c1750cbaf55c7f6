"""SM clock / power while one kernel runs back to back for ~3 s (is a kernel power-capped?)."""
import os, subprocess, sys, threading, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mrfa_b200
from mrfa_b200 import ops
dev = torch.device("cuda:0")
B, C, h, w = 64, 256, 64, 64
N = h * w
rows = ops.corr_rows_total(h, w)
a_op = torch.randn(B, rows, C, device=dev).bfloat16()
b_op = torch.randn(B, N, C, device=dev).bfloat16()
v0 = torch.empty((B, rows, N), device=dev, dtype=torch.bfloat16)
v1 = torch.empty((B, rows, N // 4), device=dev, dtype=torch.bfloat16)
st = lambda: ops._stream()
gemm = lambda: ops.check(ops.lib.mrfa_corr_volume(ops._p(a_op), ops._p(b_op), ops._p(v0), ops._p(v1), B, C, h, w, C ** -0.5, 148, st()))
big = torch.empty(1 << 30, device=dev)
kernels = {"corr_volume": gemm, "memset4GiB": lambda: big.zero_(), "bmm_bf16": lambda: torch.bmm(a_op[:, :4096], b_op.transpose(1, 2))}
rows_out = []
def sample(stop):
    while not stop.is_set():
        out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-i", "0"], capture_output=True, text=True).stdout.strip()
        rows_out.append(out)
        time.sleep(0.1)
for name, fn in kernels.items():
    for _ in range(3): fn()
    torch.cuda.synchronize()
    rows_out.clear()
    stop = threading.Event(); th = threading.Thread(target=sample, args=(stop,)); th.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time(); n = 0
    e0.record()
    while time.time() - t0 < 3.0:
        for _ in range(50): fn()
        n += 50
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    stop.set(); th.join()
    print(name, "avg ms", round(e0.elapsed_time(e1) / n, 4), "samples:", rows_out[2:-1][:12])
