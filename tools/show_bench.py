"""Pretty-print bench.py JSON lines from stdin (one per line)."""
import json
import sys

for line in sys.stdin:
    if not line.startswith("{"):
        print(line.strip()[:300])
        continue
    d = json.loads(line)
    if d.get("impl") == "reference":
        print("reference", round(d["value"], 2), d["unit"], d["cpu_baseline"]["cores"], "cores")
        continue
    nf = d["config"]["frames_per_step"]
    print(round(d["value"]), d["unit"], "ms/step", round(d["ms_per_step"], 3), "frac", round(d["step_frac_of_peak"], 3), d["clocks"])
    for k in d["kernels"]:
        print("   %-13s launches %4d  avg %8.1f us  %6.0f GB/s (%.0f%% of peak)" % (
            k["kernel"], k["launches"], k["avg_ms"] * 1000, k["achieved_gbs"], 100 * k["achieved_gbs"] / d["roofline"]["peak"]))
    print("   roofline", d["roofline"]["kernel"], round(d["roofline"]["frac"], 3), "| e2e", d["e2e"] and round(d["e2e"]["value"]), "| launches", d["gpu_launches"], "| cpu", d["cpu_baseline"] and round(d["cpu_baseline"]["value"], 1))
