#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove tcgen05 / TMEM / TMA / 256-bit accesses (B200_PROFILING.md), read from
the built library with `cuobjdump -sass`; writes profiles/sass_summary.txt.   python scripts/sass_summary.py"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "mrfa_b200", "libmrfa_b200.so")
OUT = os.path.join(ROOT, "profiles", "sass_summary.txt")
MNEMONICS = ("UTCHMMA", "UTCQMMA", "UTCBAR", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCATOMSWS", "SYNCS",
             "LDG.E.ENL2.256", "STG.E.ENL2.256", "LDG.E.128", "STG.E.128", "RED.E", "REDG", "ATOMG", "HMMA", "MATCH")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        for mn in MNEMONICS:
            if re.search(r"\b" + re.escape(mn), line):
                per[cur][mn] += 1
    names = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
    rows = []
    for (mangled, cnt), name in zip(per.items(), names):
        short = re.sub(r"\(.*", "", name).replace("void ", "")
        if sum(cnt.values()):
            rows.append((short, cnt))
    total = collections.Counter()
    for _, c in rows:
        total.update(c)
    with open(OUT, "w") as f:
        f.write("SASS evidence for libmrfa_b200.so (sm_100a), produced by scripts/sass_summary.py from `cuobjdump -sass`.\n")
        f.write("UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), UTMALDG / UTMASTG = TMA bulk tensor load / store, LDTM = tcgen05.ld,\n"
                "UTCBAR = tcgen05.commit, LDG/STG.E.ENL2.256 = 256-bit global accesses, RED = red.global.add.\n\n")
        f.write(f"{len(per)} kernels in the library; totals: " + ", ".join(f"{k} {v}" for k, v in sorted(total.items())) + "\n")
        two = sum(1 for l in sass.splitlines() if "UTCHMMA.2CTA" in l)
        f.write(f"UTCHMMA.2CTA (cta_group::2): {two}\n\n")
        for short, cnt in sorted(rows, key=lambda r: r[0]):
            f.write(f"{short}\n    " + "  ".join(f"{k}={v}" for k, v in sorted(cnt.items())) + "\n")
    print(open(OUT).read()[:3000])


if __name__ == "__main__":
    sys.exit(main())
