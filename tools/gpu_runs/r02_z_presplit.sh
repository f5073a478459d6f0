#!/bin/bash
# round 2, call Z (1 GPU): persistent rollout kernel with the pre-split (fp16 hi / lo) h1 hand-off: layer 2 = SS MMAs, no operand conversion
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_engine.py tests/test_gpu_fullsize.py -x -q -k "fused_policy_step or cfg2_4096x32 or full_size_properties" > gpurun_out/r02_z_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r02_z_pytest.log
timeout 300 python tools/rollout_trace.py > gpurun_out/r02_z_trace.log 2>&1; tail -15 gpurun_out/r02_z_trace.log
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-async > gpurun_out/r02_z_bench.log 2>&1
grep -o '"value": [0-9.]*, "unit": "env-steps/s", "n_gpus"' gpurun_out/r02_z_bench.log; grep -o '"rollout_ms": [0-9.]*' gpurun_out/r02_z_bench.log
