#!/bin/bash
# round 2, call J (1 GPU): persistent rollout kernel v2 (weight prefetch across the cluster barriers, batched 12-warp tail)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_engine.py -x -q -k "fused_policy_step or closed_loop or full_size_properties" > gpurun_out/r02_j_pytest_a.log 2>&1
echo "equivalence tests rc=$?"; tail -5 gpurun_out/r02_j_pytest_a.log | cut -c1-300
for rf in 1 0; do
  SFB200_ROLLOUT_FUSED=$rf timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-async > gpurun_out/r02_j_bench_rf$rf.log 2>&1
  echo "bench rollout_fused=$rf rc=$?"; grep -o '"value": [0-9.]*, "unit": "env-steps/s", "n_gpus"' gpurun_out/r02_j_bench_rf$rf.log; grep -o '"rollout_ms": [0-9.]*' gpurun_out/r02_j_bench_rf$rf.log
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rollout_mlp2 -s 1 -c 1 -o gpurun_out/r02_j_rollout python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-e2e --no-async > gpurun_out/r02_j_ncu.log 2>&1
echo "ncu rc=$?"
