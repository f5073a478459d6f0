#!/bin/bash
# round 2, call L (1 GPU): shared-memory address-space fix + branch-free ELU; what paces the dW GEMM (probe bits); full GPU suite
mkdir -p gpurun_out
for p in 0 2 4 6 8 14; do echo "== SFB200_TA_PROBE=$p"; SFB200_TA_PROBE=$p timeout 200 python tools/dw_bench.py 2>&1 | grep "^M="; done > gpurun_out/r02_l_dw_probe.log 2>&1; cat gpurun_out/r02_l_dw_probe.log
timeout 300 python tools/rollout_trace.py > gpurun_out/r02_l_trace.log 2>&1; tail -18 gpurun_out/r02_l_trace.log
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02_l_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02_l_pytest.log
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-async > gpurun_out/r02_l_bench.log 2>&1
grep -o '"value": [0-9.]*, "unit": "env-steps/s", "n_gpus"' gpurun_out/r02_l_bench.log; grep -o '"rollout_ms": [0-9.]*' gpurun_out/r02_l_bench.log; grep -o '"avg_kernel_ms": [0-9.]*' gpurun_out/r02_l_bench.log | head -3
