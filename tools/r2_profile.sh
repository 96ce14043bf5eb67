#!/bin/bash
# Round-2 profile capture (run on the GPU box through gpurun; every step is bounded by `timeout`).
#   1. ncu --set full of every kernel variant of the two headline workloads (tools/profile_ops.py: one round of every op)
#   2. the launch list of bench.py itself (--metrics gpu__time_duration.sum), after the same command has run clean without ncu
# Raw pages are exported to gpurun_out/ as CSV (the .ncu-rep files are too large for the return channel); tools/profile_report.py
# turns them into the tracked summaries under profiles/.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for wl in video4k:11 image1080p:37; do
  name=${wl%%:*}; batch=${wl##*:}
  timeout 120 python tools/profile_ops.py --workload $name --batch $batch --rounds 1 || exit 1
  timeout 600 ncu --set full --clock-control none --import-source on -f -o /tmp/r2_$name python tools/profile_ops.py --workload $name --batch $batch --rounds 1 > gpurun_out/r2_ncu_$name.log 2>&1
  timeout 120 ncu -i /tmp/r2_$name.ncu-rep --page raw --csv > gpurun_out/r2_${name}_raw.csv 2>/dev/null
done
timeout 300 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary > gpurun_out/r2_launchlist_bench.json 2> gpurun_out/r2_launchlist_bench.err || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2_video4k_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary > gpurun_out/r2_launchlist_ncu.log 2>&1
ls -la gpurun_out/r2_*raw.csv gpurun_out/r2_video4k_launches.csv
