#!/usr/bin/env python
"""print the stage table of bench.py JSON lines: python scripts/show_bench.py file.json ..."""
import json, sys
for f in sys.argv[1:]:
    try:
        l = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable:", e)
        continue
    w = l.get("whole_step_roofline") or {}
    print("%s: %.1f %s, %.4f ms/step (wall %.4f), SOR iters/step %s, whole-step frac %.3f, launches %s" % (
        f, l["value"], l["unit"], l["ms_per_step"], l.get("wall_ms_per_step", 0),
        l["config"].get("sor_iters_per_step"), w.get("frac", 0), l.get("gpu_launches")))
    if l.get("roofline"):
        tot = 0.0
        for k, s in l["roofline"]["stages"].items():
            tot += s["ms_per_launch"] * s["launches"] / l["steps"]
            print("    %-5s %4d launches  %.4f ms/launch  %5.1f B/pt  frac %.3f  share %.3f" % (
                k, s["launches"], s["ms_per_launch"], s["bytes_per_pt"], s["frac"], s["share_of_step"]))
        print("    stages sum %.4f ms/step -> outside spans %.4f ms" % (tot, l["ms_per_step"] - tot))
    for k in ("e2e", "e2e_resident", "cpu_baseline"):
        if l.get(k):
            print("    %s: %.1f %s" % (k, l[k]["value"], l[k].get("unit", "")))
