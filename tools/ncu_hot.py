"""Hot SASS regions of an ncu source-page CSV: python tools/ncu_hot.py file.csv [min_share]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
iS = hdr.index("# Samples"); iI = hdr.index("Instructions Executed"); iSrc = hdr.index("Source")
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = rows[2:]
tot = sum(int(r[iS]) for r in data)
toti = sum(int(r[iI]) for r in data)
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.004
print("total samples", tot, "warp instructions", toti)
for n, r in enumerate(data):
    s = int(r[iS])
    if s >= thr * tot:
        top = sorted(((int(r[i]), hdr[i][6:]) for i in stall), reverse=True)[:2]
        print(f"{n:5d} {100*s/tot:5.2f}% inst {int(r[iI]):9d} {r[iSrc].strip():60s} {top}")
