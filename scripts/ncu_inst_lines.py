"""`ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` -> executed (warp-level) instructions per source line,
with the average number of active threads, per kernel.  Issue-bound kernels: this is where the issue slots go.

usage: python scripts/ncu_inst_lines.py correlated.csv [top_n] [kernel_substring]
"""
import csv
import sys

rows = csv.reader(open(sys.argv[1]))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
want = sys.argv[3] if len(sys.argv) > 3 else ""
kernels, cur, fname, cols = {}, None, None, None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif len(r) >= 2 and r[0] == "Function Name":
        cur = kernels.setdefault(r[1].split("(")[0][-40:], [])
    elif r and r[0] == "Line No":
        cols = {c: i for i, c in enumerate(r)}
    elif cur is not None and cols and len(r) > 9 and r[0].isdigit() and r[2] == "-":
        try:
            cur.append((int(r[cols["Instructions Executed"]]), int(r[cols["Thread Instructions Executed"]]),
                        int(r[cols["# Samples"]]), fname, int(r[0]), r[1].strip()[:90]))
        except (ValueError, KeyError):
            pass
for k, agg in kernels.items():
    if want not in k:
        continue
    tot = sum(a[0] for a in agg) or 1
    tth = sum(a[1] for a in agg) or 1
    print(f"== {k}: {tot} warp instructions, {tth} thread instructions, {tth / tot:.2f} threads per instruction")
    for a in sorted(agg, reverse=True)[:top]:
        print(f"{a[0]:10d} {100 * a[0] / tot:5.1f}%  thr/inst {a[1] / max(a[0], 1):5.1f}  samples {a[2]:6d}  {a[3]}:{a[4]}  {a[5]}")
