#!/bin/bash
# ncu launch list of the final round-1 build (same command as the bench, eager so every launch is visible)
mkdir -p gpurun_out
timeout 700 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r01m.csv python bench.py --steps 2 --warmup 3 --no-graph --no-e2e --no-cpu-baseline --no-async > gpurun_out/ncu_bench43.log 2>&1; echo "ncu rc=$?"
python profiles/summarize_launches.py gpurun_out/launches_r01m.csv > gpurun_out/launches_r01m.md; head -30 gpurun_out/launches_r01m.md
