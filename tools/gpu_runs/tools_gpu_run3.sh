#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/gemm_bench.py > gpurun_out/gemm_bench.log 2>&1; echo "gemm_bench rc=$?"; cat gpurun_out/gemm_bench.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 70 -c 3 -o gpurun_out/prof_gemm_tc python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-graph > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:heads_forward_kernel -s 40 -c 2 -o gpurun_out/prof_heads_fwd python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-graph > gpurun_out/ncu_full2.log 2>&1; echo "ncu heads rc=$?"
ls -la gpurun_out
