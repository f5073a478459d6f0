#!/bin/bash
# round 2, call K (1 GPU): phase trace of the persistent rollout kernel + bench with the cheaper proxy fence
mkdir -p gpurun_out
timeout 300 python tools/rollout_trace.py > gpurun_out/r02_k_trace.log 2>&1; tail -14 gpurun_out/r02_k_trace.log
timeout 200 python -m pytest tests/test_gpu_engine.py -x -q -k "fused_policy_step" > gpurun_out/r02_k_pytest.log 2>&1; echo "equiv rc=$?"; tail -2 gpurun_out/r02_k_pytest.log
SFB200_ROLLOUT_FUSED=1 timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-async > gpurun_out/r02_k_bench_rf1.log 2>&1
grep -o '"value": [0-9.]*, "unit": "env-steps/s", "n_gpus"' gpurun_out/r02_k_bench_rf1.log; grep -o '"rollout_ms": [0-9.]*' gpurun_out/r02_k_bench_rf1.log
