"""print the headline numbers of a bench.py JSON line (helper for reading gpurun_out/*.json)"""
import json
import sys

d = json.load(open(sys.argv[1]))
print("value %.4g  %.3f ms/step  roofline %.3f (%s)  e2e %.4g  cpu %s  launches %s" % (
    d["value"], d["ms_per_step"], d["roofline"]["frac"] or 0, d["roofline"]["kernel"][:40], (d.get("e2e") or {}).get("value", 0),
    (d.get("cpu_baseline") or {}).get("value"), d.get("gpu_launches")))
print("  sections", {a: round(b, 3) for a, b in d["roofline"]["sections_ms_per_step"].items()})
print("  config", {k: v for k, v in d["config"].items() if k not in ("workload", "l2")})
for k, v in d.get("other_configs", {}).items():
    print(k, "%.3g /s" % v["value"], "%.2f ms" % v["ms_per_step"], "n=%d" % v["particles"], v["roofline"]["kernel"][:28], "frac %.3f" % (v["roofline"]["frac"] or 0))
    print("    ", {a: round(b, 3) for a, b in v["sections_ms_per_step"].items()}, v.get("sort_path"), "extras", v.get("sort_extras_last_step"),
          v.get("weight_conservation"), v.get("initial_merge"), v.get("merges_in_timed_steps"), "T=%.2f" % (v.get("mean_T_K") or 0))
