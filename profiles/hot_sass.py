"""Lists the hot SASS instructions of an `ncu --page source --csv` export (share of executed instructions),
plus per-opcode totals (instructions, shared-memory wavefronts)."""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.002
rows = list(csv.reader(open(path)))
hdr = rows[1]
col = {n: hdr.index(n) for n in ("Address", "Source", "Instructions Executed", "# Samples", "L1 Wavefronts Shared", "L1 Wavefronts Shared Ideal")}
data = []
for r in rows[2:]:
    if r and r[0] == "Kernel Name":
        break
    try:
        data.append((r[col["Address"]], r[col["Source"]], int(r[col["Instructions Executed"]]), int(r[col["# Samples"]]),
                     int(r[col["L1 Wavefronts Shared"]] or 0), int(r[col["L1 Wavefronts Shared Ideal"]] or 0)))
    except Exception:
        pass
tot = sum(d[2] for d in data)
smp = sum(d[3] for d in data)
print("total warp instructions", tot, "sass lines", len(data), "samples", smp)
ops = defaultdict(lambda: [0, 0, 0, 0])
for d in data:
    op = d[1].split()
    op = op[1] if op and op[0].startswith("@") else (op[0] if op else "?")
    op = ".".join(op.split(".")[:2]) if op.startswith(("LDS", "STS", "LDG", "STG", "LDL", "STL")) else op.split(".")[0]
    o = ops[op]
    o[0] += d[2]; o[1] += d[4]; o[2] += d[5]; o[3] += d[3]
print("opcode            inst%   samples%  wavefronts(shared)  ideal")
for op, o in sorted(ops.items(), key=lambda kv: -kv[1][0])[:28]:
    print(f"{op:16s} {o[0] / tot * 100:6.2f}  {o[3] / max(smp, 1) * 100:6.2f}  {o[1]:14d} {o[2]:14d}")
if "-l" in sys.argv:
    for i, d in enumerate(data):
        if d[2] > tot * thr:
            print(i, d[0][-5:], f"{d[2] / tot * 100:5.2f}% smp{d[3]:6d} wf{d[4]:10d}", d[1][:110])
