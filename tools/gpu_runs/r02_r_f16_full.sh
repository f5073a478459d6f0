#!/bin/bash
# round 2, call R (1 GPU): fp16-split engine wired into sampler (per-step + persistent) and learner: full GPU suite + bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02_r_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r02_r_pytest.log
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-async > gpurun_out/r02_r_bench.log 2>&1
grep -o '"value": [0-9.]*, "unit": "env-steps/s", "n_gpus"' gpurun_out/r02_r_bench.log; grep -o '"rollout_ms": [0-9.]*' gpurun_out/r02_r_bench.log; grep -o '"avg_kernel_ms": [0-9.]*' gpurun_out/r02_r_bench.log | head -3; tail -3 gpurun_out/r02_r_bench.log | cut -c1-300
timeout 300 python tools/rollout_trace.py > gpurun_out/r02_r_trace.log 2>&1; tail -16 gpurun_out/r02_r_trace.log
