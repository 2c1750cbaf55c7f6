"""Condense an ncu launch list (`--metrics gpu__time_duration.sum --csv --log-file launches.csv`) into the per-kernel
table kept under profiles/.

    python scripts/launch_summary.py launches.csv out.md "title"
"""
import csv
import re
import sys
from collections import OrderedDict

src, out_md = sys.argv[1:3]
title = sys.argv[3] if len(sys.argv) > 3 else "ncu launch list of one refinement step"
rows = list(csv.reader(open(src, newline="")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
names = rows[hdr]
kn, mv, mu = names.index("Kernel Name"), names.index("Metric Value"), names.index("Metric Unit")
agg = OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= mv or "gpu__time_duration" not in r[names.index("Metric Name")]:
        continue
    v = float(r[mv].replace(",", ""))
    us = v * {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}.get(r[mu], 1e-3)
    name = re.sub(r"^void ", "", r[kn])
    name = re.sub(r"\(.*", "", name)[:110]
    a = agg.setdefault(name, [0.0, 0])
    a[0] += us
    a[1] += 1
total = sum(a[0] for a in agg.values())
n = sum(a[1] for a in agg.values())
ours = sum(a[0] for k, a in agg.items() if k.startswith("mrfa::"))
n_ours = sum(a[1] for k, a in agg.items() if k.startswith("mrfa::"))
aten = sum(a[0] for k, a in agg.items() if k.startswith("at::"))
lines = [f"# {title}", "",
         "`ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none` over `scripts/profile_step.py`.",
         "Per-launch times are cold-cache and serialised; the SHARES are what is compared with `bench.py`'s in-step timings.", "",
         f"Total {total / 1e3:.2f} ms in {n} launches: this library {ours / 1e3:.2f} ms ({n_ours} launches, {100 * ours / total:.1f} %), "
         f"ATen elementwise / layout {aten / 1e3:.2f} ms ({100 * aten / total:.1f} %), cuDNN / cuBLAS / other "
         f"{(total - ours - aten) / 1e3:.2f} ms ({100 * (total - ours - aten) / total:.1f} %).", "",
         "| us | launches | share % | kernel |", "|---:|---:|---:|---|"]
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    lines.append(f"| {a[0]:.1f} | {a[1]} | {100 * a[0] / total:.2f} | `{k}` |")
open(out_md, "w").write("\n".join(lines) + "\n")
print(lines[5])
