#!/bin/bash
# round 2, call E (1 GPU): fused policy kernel v2 (16 conversion warps) + fused tail; A/B timings; dW operand-warp A/B; full suite
mkdir -p gpurun_out
REPS=200 timeout 120 python tools/ncu_policy_step.py > gpurun_out/r02_e_ps_timing.log 2>&1; tail -1 gpurun_out/r02_e_ps_timing.log
timeout 300 python -m pytest tests/test_gpu_engine.py tests/test_gpu_kernels.py -x -q -k "fused_policy_step or policy_mlp2 or shuffle or closed_loop" > gpurun_out/r02_e_pytest_a.log 2>&1
echo "targeted tests rc=$?"; tail -6 gpurun_out/r02_e_pytest_a.log
timeout 120 python tools/dw_bench.py > gpurun_out/r02_e_dw_opw4.log 2>&1; grep "M=32768 N=512 K=512" gpurun_out/r02_e_dw_opw4.log
SFB200_TA_DW_OPW=8 timeout 120 python tools/dw_bench.py > gpurun_out/r02_e_dw_opw8.log 2>&1; grep "M=32768 N=512 K=512" gpurun_out/r02_e_dw_opw8.log
for mode in "1 1" "0 1" "0 0"; do
  set -- $mode
  SFB200_POLICY_FUSED=$1 SFB200_TAIL_FUSED=$2 timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-async > gpurun_out/r02_e_bench_pf$1_tf$2.log 2>&1
  echo "bench policy_fused=$1 tail_fused=$2 rc=$?"; grep -o '"value": [0-9.]*, "unit": "env-steps/s", "n_gpus"' gpurun_out/r02_e_bench_pf$1_tf$2.log; grep -o '"rollout_ms": [0-9.]*' gpurun_out/r02_e_bench_pf$1_tf$2.log; grep -o '"launches_per_step": {[^}]*}' gpurun_out/r02_e_bench_pf$1_tf$2.log
done
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_e_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/r02_e_pytest_gpu.log
