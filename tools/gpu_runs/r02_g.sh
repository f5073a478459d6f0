#!/bin/bash
# round 2, call G (1 GPU): optimized fused tail A/B, conv gradient accuracy by engine, targeted tests
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_engine.py -x -q -k "fused_policy_step or shuffle or closed_loop or golden and cfg2_small" > gpurun_out/r02_g_pytest_a.log 2>&1
echo "targeted tests rc=$?"; tail -4 gpurun_out/r02_g_pytest_a.log
timeout 300 python tools/conv_grad_check.py > gpurun_out/r02_g_conv_grad.log 2>&1; cat gpurun_out/r02_g_conv_grad.log | tail -30
for tf in 1 0; do
  SFB200_TAIL_FUSED=$tf timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-async > gpurun_out/r02_g_bench_tf$tf.log 2>&1
  echo "bench tail_fused=$tf rc=$?"; grep -o '"value": [0-9.]*, "unit": "env-steps/s", "n_gpus"' gpurun_out/r02_g_bench_tf$tf.log; grep -o '"rollout_ms": [0-9.]*' gpurun_out/r02_g_bench_tf$tf.log
done
