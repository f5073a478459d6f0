#!/bin/bash
# round-1 GPU session 2: validate the tcgen05 engine, then full tests, bench (both engines), ncu launch list
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -k "tc_engine or 3xtf32 or heads_backward" -s --timeout 120 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_tc.log 2>&1
TC_RC=$?
echo "tc tests rc=$TC_RC"; tail -15 gpurun_out/pytest_tc.log
if [ $TC_RC -eq 0 ]; then
  timeout 900 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
  echo "full gpu tests rc=$?"; tail -5 gpurun_out/pytest_gpu.log
  timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_tc.log 2>&1; echo "bench tc rc=$?"; tail -1 gpurun_out/bench_tc.log
  timeout 300 python bench.py --steps 10 --warmup 3 --engine simt --no-e2e --no-cpu-baseline > gpurun_out/bench_simt.log 2>&1; echo "bench simt rc=$?"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_tc.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-graph > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
else
  timeout 900 python -m pytest tests -m gpu -q -k "not 3xtf32 and not tc_engine" --timeout 300 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
  echo "gpu tests (no tc) rc=$?"; tail -5 gpurun_out/pytest_gpu.log
fi
nvidia-smi --query-gpu=name,temperature.gpu,clocks.sm --format=csv
