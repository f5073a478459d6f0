#!/bin/bash
# re-entry baseline: full gpu tests, bench (all arms), ncu launch list of the current build
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/all_tests.log 2>&1; echo "all tests rc=$?"; tail -4 gpurun_out/all_tests.log
timeout 300 python tools/gemm_bench.py > gpurun_out/gemm_bench.log 2>&1; grep -E "3xtf32|heads" gpurun_out/gemm_bench.log
timeout 900 python bench.py > gpurun_out/bench13.json 2> gpurun_out/bench13.err; echo "bench rc=$?"; tail -3 gpurun_out/bench13.err; cat gpurun_out/bench13.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r01e.csv python bench.py --steps 2 --warmup 3 --no-graph --no-e2e --no-cpu-baseline --no-async > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
