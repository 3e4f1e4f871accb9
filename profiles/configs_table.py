"""Markdown table of the `configs` block of a bench.py line (the five BASELINE.json configurations).  usage: python profiles/configs_table.py bench.json"""
import json
import sys

d = json.load(open(sys.argv[1]))
print("| configuration | workload | kernel shape | frames | device ms | coded Gb/s | G edge-it/s | roofline (bound: achieved / peak = frac) | FP64 pipe |")
print("|---|---|---|---|---|---|---|---|---|")
rows = [("headline", {"workload": d["config"]["workload"], "kernel_shape": {k: d["config"].get(k) for k in ("frames_per_cta", "threads_per_cta", "ctas", "residency")},
                      "frames": d["config"]["frames_per_step_per_gpu"], "device_ms": d["roofline"]["kernel_ms"], "value": d["value"],
                      "edge_updates_per_s": d["edge_updates_per_s"], "roofline": d["roofline"]})] + list(d.get("configs", {}).items())
for name, c in rows:
    ks, r, f = c["kernel_shape"], c["roofline"], c.get("fp64_pipe")
    shape = "%s frames/CTA x %s threads x %s CTAs, %s" % (ks["frames_per_cta"], ks["threads_per_cta"], ks["ctas"], ks["residency"])
    roof = "%s: %.0f / %.0f GB/s = **%.2f**" % (r["bound"], r["achieved"], r["peak"], r["frac"])
    if r.get("traffic"):
        roof += "; DRAM %.3g B per launch" % r["traffic"]
    fp = "%.2f (%.1f instr per edge-it)" % (f["frac"], f["fp64_instructions_per_edge_iteration"]) if f else ""
    print("| %s | %s | %s | %d | %.1f | %.3f | %.1f | %s | %s |" % (name, c["workload"], shape, c["frames"], c["device_ms"], c["value"], c["edge_updates_per_s"] / 1e9, roof, fp))
for k in ("e2e", "e2e_f64", "e2e_simulate", "et_on", "f32_messages", "decode_single_frame", "cpu_baseline"):
    if k in d:
        print("\n`%s`: %s" % (k, json.dumps(d[k])))
