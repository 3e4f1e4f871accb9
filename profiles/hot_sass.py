"""Lists the hot SASS instructions of an `ncu --page source --csv` export (share of executed instructions)."""
import csv
import sys

path = sys.argv[1]
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.002
rows = list(csv.reader(open(path)))
hdr = rows[1]
ia, isrc, iex, ismp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
data = []
for r in rows[2:]:
    if r and r[0] == "Kernel Name":
        break
    try:
        data.append((r[ia], r[isrc], int(r[iex]), int(r[ismp])))
    except Exception:
        pass
tot = sum(d[2] for d in data)
print("total warp instructions", tot, "sass lines", len(data))
for i, d in enumerate(data):
    if d[2] > tot * thr:
        print(i, d[0][-5:], f"{d[2] / tot * 100:5.2f}% smp{d[3]:6d}", d[1][:120])
