#!/usr/bin/env python
"""Rank CUDA source lines of an `ncu --page source --csv --print-source cuda,sass` dump by a per-line metric.
Usage: ncu_lines.py dump.csv ["L1 Wavefronts Shared" | "# Samples" | ...] [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
metric = sys.argv[2] if len(sys.argv) > 2 else "L1 Wavefronts Shared"
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
hdr = next(r for r in rows if r and r[0] == "Line No")
col = hdr.index(metric)
ideal = hdr.index(metric + " Ideal") if metric + " Ideal" in hdr else None
def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return 0.0
out, cur = [], None
for r in rows:
    if r and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif len(r) > col and r[0].strip().isdigit():
        v = num(r[col])
        if v > 0:
            out.append((v, num(r[ideal]) if ideal is not None else 0.0, cur, r[0], r[1].strip()[:100]))
tot = sum(o[0] for o in out)
print(f"total {metric}: {tot:.4g}")
for v, i, fn, ln, src in sorted(out, reverse=True)[:top]:
    print(f"{100 * v / tot:5.1f}%  {v:.3g}" + (f" (ideal {i:.3g})" if ideal is not None else "") + f"  {fn}:{ln}  {src}")
