"""Stall samples of an ncu source page aggregated by CUDA source line.
ncu -i rep --page source --print-source cuda,sass --csv > f.csv; python tools/ncu_lines.py f.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur = None; agg = []; hdr = None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split('/')[-1]; continue
    if len(r) > 5 and r[0] == "Line No":
        hdr = r; iS = r.index("# Samples"); iI = r.index("Instructions Executed"); continue
    if hdr and len(r) > iI and r[0].isdigit() and r[2] == "-":
        try:
            agg.append((int(r[iS]), int(r[iI]), cur, int(r[0]), r[1].strip()[:120]))
        except ValueError:
            pass
tot = sum(a[0] for a in agg); toti = sum(a[1] for a in agg)
print("total samples", tot, "warp instructions", toti)
for a in sorted(agg, reverse=True)[:top]:
    print(f"{100 * a[0] / tot:5.2f}% inst {100 * a[1] / toti:5.2f}%  {a[2]}:{a[3]}  {a[4]}")
