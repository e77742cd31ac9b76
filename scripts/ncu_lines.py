"""`ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` -> warp-stall samples per source line, per kernel.

usage: python scripts/ncu_lines.py correlated.csv [top_n]
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
kernels, cur, fname = {}, None, None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif len(r) >= 2 and r[0] == "Function Name":
        cur = kernels.setdefault(r[1].split("(")[0][-40:], [])
    elif cur is not None and len(r) >= 5 and r[0].isdigit() and r[2] == "-":
        try:
            cur.append((int(r[4]), fname, int(r[0]), r[1].strip()[:100]))
        except ValueError:
            pass
for k, agg in kernels.items():
    tot = sum(a[0] for a in agg) or 1
    print(f"== {k}: {tot} samples")
    for a in sorted(agg, reverse=True)[:top]:
        print(f"{a[0]:6d} {100 * a[0] / tot:5.1f}%  {a[1]}:{a[2]}  {a[3]}")
