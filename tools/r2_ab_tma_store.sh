#!/bin/bash
# A/B on ONE box: apply-kernel output through per-thread vector stores (shipped) vs TMA stores (WM_OPT_TMA_STORE), both headline workloads
set -u
cd "$(dirname "$0")/.."
F="--no-e2e --no-cpu-baseline --no-secondary --no-sync-proto --steps 10 --warmup 3"
for rep in 1 2; do
  for wl in video4k image1080p image4k; do
    timeout 200 python bench.py --workload $wl $F --no-tma-store > gpurun_out/ab_${wl}_stg_$rep.json 2>/dev/null
    timeout 200 python bench.py --workload $wl $F > gpurun_out/ab_${wl}_tma_$rep.json 2>/dev/null
  done
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/ab_*.json")):
    try:
        d = json.load(open(f))
    except Exception as e:
        print(f, "unreadable", e); continue
    ks = {k["kernel"]: k["avg_ms"] * 1e3 for k in d["kernels"]}
    print("%-40s %9.0f frames/s  me_apply %8.1f us  nvf_apply %8.1f us" % (f.split("/")[-1], d["value"], ks.get("me_apply", 0), ks.get("nvf_apply", 0)))
PY
