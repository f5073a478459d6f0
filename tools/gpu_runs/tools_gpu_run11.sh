#!/bin/bash
mkdir -p gpurun_out
echo "== raw_hi=1"; SFB200_TC_RAW_HI=1 timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "linear or tc_engine" -s 2>&1 | grep -E "max abs error|passed|failed|Error" | head
SFB200_TC_RAW_HI=1 timeout 300 python tools/gemm_bench.py 2>&1 | grep -E "3xtf32" 
echo "== raw_hi=0"; timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "tc_engine" -s 2>&1 | grep -E "max abs error|passed|failed" | head
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r01e.csv python bench.py --steps 2 --warmup 3 --no-graph --no-e2e --no-cpu-baseline --no-async > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
