#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 66 -c 2 -o gpurun_out/prof_gemm_tc_v2 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-graph > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/*.ncu-rep
