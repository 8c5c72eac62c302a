"""Developer aid: one line per workload of a bench.py JSON line (python scripts/show_bench.py gpurun_out/x.json)."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("headline %-8s value %.1f GTEPS  %.3f ms/step  e2e %.2f (%.1f ms)  roofline %.3f (%.0f GB/s, %.3f ms/launch)  parity %s  wall %.0fs" % (
    d["config"]["workload"][:8], d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["frac"],
    d["roofline"]["achieved"], d["roofline"]["ms_per_launch"], d.get("parity_check", {}).get("status"), d.get("wall_s", 0)))
if d.get("cpu_baseline"):
    print("  cpu:", d["cpu_baseline"].get("value"), d["cpu_baseline"].get("sample"))
for k, v in (d.get("extras") or {}).get("workloads", {}).items():
    if "value" not in v:
        print("  %-16s %s" % (k, v))
        continue
    ps = v.get("per_source_gteps") or {}
    print("  %-16s value %.1f GTEPS  %.3f ms/step  e2e %.2f (%.1f ms)  roofline %.3f (%.0f GB/s)  min/max %s/%s  iters %.1f%s" % (
        k, v["value"], v["ms_per_step"], v["e2e"]["value"], v["e2e"]["ms_per_step"], v["roofline"]["frac"], v["roofline"]["achieved"],
        ("%.0f" % ps["min"]) if ps else "-", ("%.0f" % ps["max"]) if ps else "-", v["iterations_per_run"],
        ("  cpu %.2f" % v["cpu_baseline"]["value"]) if v.get("cpu_baseline") and v["cpu_baseline"].get("value") else ""))
