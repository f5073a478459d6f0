#!/bin/bash
# round 2, call I (1 GPU): persistent whole-rollout kernel -- equivalence test (bounded), then bench A/B and the suite
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_engine.py -x -q -k "fused_policy_step" > gpurun_out/r02_i_pytest_a.log 2>&1
echo "equivalence tests rc=$?"; tail -15 gpurun_out/r02_i_pytest_a.log | cut -c1-300
for rf in 1 0; do
  SFB200_ROLLOUT_FUSED=$rf timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-async > gpurun_out/r02_i_bench_rf$rf.log 2>&1
  echo "bench rollout_fused=$rf rc=$?"; grep -o '"value": [0-9.]*, "unit": "env-steps/s", "n_gpus"' gpurun_out/r02_i_bench_rf$rf.log; grep -o '"rollout_ms": [0-9.]*' gpurun_out/r02_i_bench_rf$rf.log; grep -o '"launches_per_step": {[^}]*}' gpurun_out/r02_i_bench_rf$rf.log; tail -3 gpurun_out/r02_i_bench_rf$rf.log | cut -c1-300 | grep -i "error\|trap\|illegal"
done
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02_i_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/r02_i_pytest_gpu.log | cut -c1-300
