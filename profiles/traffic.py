"""Runs profiles/traffic_cmd.py under ncu for every bench.py configuration and writes profiles/roofline_traffic.json:
DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of ONE launch of each configuration's kernel — bench.py's
`roofline.traffic`.  usage (GPU box): python profiles/traffic.py [outdir]"""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out_dir = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out")
names = ["headline", "C1_bp_fixed50", "C1_et_sweep", "C2_et", "C5_bsc", "C5_bec", "C3_bg1_ms", "C4_dvbs2_bp_noet"]
res, detail = {}, {}
env = dict(os.environ, LDPC_B200_PAIR="1")
for n in names:
    cmd = ["ncu", "--nvtx", "--nvtx-include", "measure/", "--clock-control", "none", "--metrics",
           "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed_pipe_fp64.sum,smsp__inst_executed.sum", "--csv", sys.executable, os.path.join(ROOT, "profiles", "traffic_cmd.py"), n]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env)
    rows = [row for row in csv.reader(io.StringIO(r.stdout)) if len(row) > 10]
    open(os.path.join(out_dir, "traffic_%s.csv" % n), "w").write(r.stdout)
    while rows and "Kernel Name" not in rows[0]:
        rows.pop(0)
    if len(rows) < 2:
        detail[n] = {"error": (r.stdout + r.stderr)[-400:]}
        continue
    hdr = rows[0]
    ki, mi, ui, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    tot, ms, kern, fp64, inst = 0.0, None, None, None, None
    for row in rows[1:]:
        try:
            v = float(row[vi].replace(",", ""))
        except ValueError:
            continue
        if row[mi].startswith("dram__bytes"):
            tot += v * scale.get(row[ui], 1)
        elif row[mi].startswith("gpu__time_duration"):
            ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1, "second": 1e3}.get(row[ui], 1)
        elif row[mi] == "smsp__inst_executed_pipe_fp64.sum":
            fp64 = v
        elif row[mi] == "smsp__inst_executed.sum":
            inst = v
        kern = row[ki]
    res[n] = int(tot)
    m = re.findall(r"'edge_iterations': (\d+)", r.stdout)   # ctx.stats() of the measured launch, printed by traffic_cmd.py
    detail[n] = {"kernel": kern, "gpu_time_ms_under_ncu": ms, "dram_bytes": int(tot), "warp_instructions": inst, "fp64_warp_instructions": fp64,
                 "edge_iterations": int(m[-1]) if m else None}
    print(n, detail[n], flush=True)
# derived: FP64 thread instructions per edge-iteration of the sum-product kernels, DRAM bytes over the algorithmic 32 B per edge-iteration
res["_fp64_thread_instructions_per_edge_iteration"] = {n: 32.0 * d["fp64_warp_instructions"] / d["edge_iterations"] for n, d in detail.items()
                                                       if n in ("C1_bp_fixed50", "C1_et_sweep", "C4_dvbs2_bp_noet") and d.get("edge_iterations") and d.get("fp64_warp_instructions")}
res["_algorithmic_bytes"] = {n: 32 * d["edge_iterations"] for n, d in detail.items() if n in ("headline", "C3_bg1_ms", "C4_dvbs2_bp_noet") and d.get("edge_iterations")}
res["_dram_over_algorithmic"] = {n: detail[n]["dram_bytes"] / b for n, b in res["_algorithmic_bytes"].items()}
res["_source"] = "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum of one launch per configuration (profiles/traffic.py -> traffic_cmd.py, same frame counts as bench.py)"
res["_detail"] = detail
json.dump(res, open(os.path.join(out_dir, "roofline_traffic.json"), "w"), indent=1)
