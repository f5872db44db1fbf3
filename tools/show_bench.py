#!/usr/bin/env python
"""Pretty-print bench.py JSON lines found in the given log files."""
import json
import sys

for path in sys.argv[1:]:
    print(path)
    try:
        line = [x for x in open(path) if x.startswith("{")][-1]
        d = json.loads(line)
    except Exception as e:  # noqa: BLE001
        print("  (no result)", e)
        try:
            print(open(path).read()[-800:])
        except Exception:
            pass
        continue
    print("  value %.1f utt/s  %.2f ms/step  e2e %.1f utt/s  launches %.0f  loss %.4f  clocks %s" % (
        d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], d.get("final_loss", float("nan")), d.get("clocks")))
    r = d.get("roofline", {})
    print("  roofline: %s %.1f / %.1f %s = %.3f (share of step %.2f)" % (r.get("kernel"), r.get("achieved", 0), r.get("peak", 0), r.get("unit"), r.get("frac", 0), r.get("share_of_step", 0)))
    for g in d.get("kernels", [])[:24]:
        print("   %-24s n=%5.1f %8.2f ms %5.1f%% %7.1f TF/s" % (g["kernel"], g["launches_per_step"], g["ms_per_step"], 100 * g["share"], g["tflops"]))
    if "cpu_baseline" in d:
        print("  cpu_baseline", d["cpu_baseline"])
