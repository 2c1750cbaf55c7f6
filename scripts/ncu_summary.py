"""Condense an `ncu --set full` report into the table kept under profiles/ and the per-kernel DRAM traffic that
bench.py reports as roofline.traffic.

    ncu -i rep.ncu-rep --page raw --csv > raw.csv
    python scripts/ncu_summary.py raw.csv out.md traffic.json
"""
import csv
import json
import re
import sys
from collections import OrderedDict

raw, out_md, out_json = sys.argv[1:4]
rows = list(csv.reader(open(raw, newline="")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
names, units, data = rows[hdr], rows[hdr + 1], rows[hdr + 2:]


def col(*keys):
    for k in keys:
        for i, n in enumerate(names):
            if n == k:
                return i
    return None


C = {"name": col("Kernel Name"), "grid": col("Grid Size"), "time": col("gpu__time_duration.sum"),
     "rd": col("dram__bytes_read.sum"), "wr": col("dram__bytes_write.sum"),
     "dram": col("dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
     "l1": col("l1tex__t_sector_hit_rate.pct"), "l2": col("lts__t_sector_hit_rate.pct"),
     "occ": col("sm__warps_active.avg.pct_of_peak_sustained_active"), "regs": col("launch__registers_per_thread"),
     "sm": col("sm__throughput.avg.pct_of_peak_sustained_elapsed"),
     "tensor": col("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active",
                   "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")}


def num(r, key, scale_from_unit=None):
    i = C[key]
    if i is None or i >= len(r) or r[i] in ("", "n/a"):
        return float("nan")
    v = float(r[i].replace(",", ""))
    u = units[i].lower()
    if scale_from_unit == "ms":
        v *= {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}.get(u, 1.0)
    if scale_from_unit == "MB":
        v *= {"byte": 1e-6, "kbyte": 1e-3, "mbyte": 1.0, "gbyte": 1e3}.get(u, 1e-6)
    return v


def short(n):
    n = re.sub(r"\(.*", "", n)
    return n.replace("void ", "").replace("mrfa::", "")


traffic = OrderedDict()
lines = ["| kernel | grid | ms | DRAM read MB | DRAM write MB | DRAM % | L1 hit % | L2 hit % | occupancy % | regs | SM % | tensor % |",
         "|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|"]
for r in data:
    if len(r) <= C["name"]:
        continue
    grid = r[C["grid"]].replace(" ", "").replace(",1,1)", ")").strip("()") if C["grid"] is not None else ""
    rd, wr = num(r, "rd", "MB"), num(r, "wr", "MB")
    lines.append("| `%s` | %s | %.4f | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %d | %.1f | %.1f |" % (
        short(r[C["name"]]), grid, num(r, "time", "ms"), rd, wr, num(r, "dram"), num(r, "l1"), num(r, "l2"), num(r, "occ"),
        int(num(r, "regs")) if num(r, "regs") == num(r, "regs") else 0, num(r, "sm"), num(r, "tensor")))
    key = re.sub(r"(_nhwc|_nchw)?(_run|_q|_v4|_scalar|_tma|_2sm|_tiled|_tiled_rows|_s4k13|_fewc|_finish)*_kernel.*", "", short(r[C["name"]]).split("<")[0])
    # the names bench.py's KernelTimer uses (ops.py `_timed`): one entry per C-ABI call, whatever kernels it launches
    key = {"cast_bf16": "corr_pack", "avg_pool2x2": "avg_pool2x2_nhwc" if "nhwc" in r[C["name"]] else "avg_pool2x2",
           "occlusion_blend_subpixel": "occlusion_blend"}.get(key, key)
    traffic[key] = traffic.get(key, 0.0) + (rd + wr) * 1e6
open(out_md, "w").write("\n".join(lines) + "\n")
json.dump({k: round(v) for k, v in traffic.items()}, open(out_json, "w"), indent=1)
print(len(lines) - 2, "kernel launches summarised;", ", ".join(f"{k}={v / 1e9:.2f}GB" for k, v in traffic.items()))
