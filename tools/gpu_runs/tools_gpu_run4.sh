#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -k "tc_engine or 3xtf32 or heads_backward or sampler_pre_post or moments" -s --timeout 120 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_tc.log 2>&1
TC_RC=$?
echo "tc tests rc=$TC_RC"; tail -8 gpurun_out/pytest_tc.log
timeout 300 python tools/gemm_bench.py > gpurun_out/gemm_bench.log 2>&1; echo "gemm_bench rc=$?"; cat gpurun_out/gemm_bench.log
if [ $TC_RC -eq 0 ]; then
  timeout 900 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
  echo "full gpu tests rc=$?"; tail -5 gpurun_out/pytest_gpu.log
  timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_tc.log 2>&1; echo "bench tc rc=$?"; tail -1 gpurun_out/bench_tc.log | cut -c1-300
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_tc.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-graph > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
fi
