#!/bin/bash
# round 2, call D (1 GPU): fused step tail + fused policy step: equivalence tests, full suite, bench A/B, ncu of the fused GEMM
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_engine.py -x -q -k "fused_policy_step or golden and cfg2_small or closed_loop" > gpurun_out/r02_d_pytest_tail.log 2>&1
echo "tail tests rc=$?"; tail -12 gpurun_out/r02_d_pytest_tail.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:policy_mlp2 -s 2 -c 2 -o gpurun_out/r02_d_policy_step python tools/ncu_policy_step.py > gpurun_out/r02_d_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/r02_d_ncu.log
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-async > gpurun_out/r02_d_bench_n1.log 2>&1
echo "bench rc=$?"; grep -o '"value": [0-9.]*, "unit": "env-steps/s", "n_gpus"' gpurun_out/r02_d_bench_n1.log; grep -o '"rollout_ms": [0-9.]*' gpurun_out/r02_d_bench_n1.log; grep -o '"launches_per_step": {[^}]*}' gpurun_out/r02_d_bench_n1.log
SFB200_POLICY_FUSED=0 timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-async > gpurun_out/r02_d_bench_n1_gemm_unfused.log 2>&1
echo "bench (per-layer GEMMs, fused tail) rc=$?"; grep -o '"value": [0-9.]*, "unit": "env-steps/s", "n_gpus"' gpurun_out/r02_d_bench_n1_gemm_unfused.log; grep -o '"rollout_ms": [0-9.]*' gpurun_out/r02_d_bench_n1_gemm_unfused.log
timeout 120 python tools/dw_bench.py > gpurun_out/r02_d_dw_bench.log 2>&1; cat gpurun_out/r02_d_dw_bench.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_d_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/r02_d_pytest_gpu.log
