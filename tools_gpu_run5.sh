#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -k "heads" --timeout 120 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_heads.log 2>&1; echo "heads tests rc=$?"; tail -3 gpurun_out/pytest_heads.log
timeout 300 python tools/gemm_bench.py 2>&1 | grep heads
# source-level capture of the learner L1 forward GEMM (M=32768, K=64: epilogue-dominated) and L2 forward (K=512)
timeout 600 ncu --set full --clock-control none --import-source on --source-level sass -k regex:gemm_tc_kernel -s 66 -c 2 -o gpurun_out/prof_gemm_tc_v2 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-graph > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/*.ncu-rep
