"""Aggregate an `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` export by CUDA source line.

    ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass > src.csv
    python scripts/ncu_source_hot.py src.csv [top_n]

Prints, per source file, the lines with the most warp-stall samples next to their executed warp instructions.
"""
import csv
import sys


def num(x):
    try:
        return float(x)
    except ValueError:
        return 0.0


def main(path, top_n=40):
    rows = list(csv.reader(open(path)))
    sections, cur = [], None
    for r in rows:
        if r and r[0] == "File Path":
            cur = {"file": r[1], "rows": []}
            sections.append(cur)
        elif r and r[0] == "Line No":
            cur["hdr"] = r
        elif r and r[0] == "Function Name":
            continue
        elif cur is not None and "hdr" in cur:
            cur["rows"].append(r)
    for s in sections:
        h = s["hdr"]
        iL, iS, iI, iSm = h.index("Line No"), h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
        tot = sum(num(r[iI]) for r in s["rows"])
        tots = sum(num(r[iSm]) for r in s["rows"])
        print(f"== {s['file']}: {len(s['rows'])} rows, {tot:.0f} warp instructions, {tots:.0f} samples")
        for r in sorted(s["rows"], key=lambda r: -num(r[iSm]))[:top_n]:
            print(f"{r[iL]:>6} inst={num(r[iI]):>10.0f} samples={num(r[iSm]):>7.0f}  {r[iS][:120]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
