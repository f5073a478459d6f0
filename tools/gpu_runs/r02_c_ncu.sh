#!/bin/bash
# round 2, call C (1 GPU): ncu --set full of the fused policy step vs the per-layer GEMMs, event timings, re-run of the two
# tests that failed in call B, first runs of bench.py --config 3/4/5
mkdir -p gpurun_out
REPS=200 timeout 120 python tools/ncu_policy_step.py > gpurun_out/r02_c_ps_timing.log 2>&1; tail -2 gpurun_out/r02_c_ps_timing.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"policy_mlp2|gemm_tc_ta" -s 8 -c 6 -o gpurun_out/r02_c_policy_step python tools/ncu_policy_step.py > gpurun_out/r02_c_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r02_c_ncu.log
timeout 900 python -m pytest tests/test_boundary.py tests/test_gpu_fullsize.py -m gpu -q -s > gpurun_out/r02_c_pytest.log 2>&1
echo "pytest rc=$?"; grep -E "CartPole running|passed|failed|near-tie" gpurun_out/r02_c_pytest.log | tail -8
for c in 3 4 5; do
  timeout 600 python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/r02_c_bench_cfg$c.log 2>&1
  echo "bench cfg$c rc=$?"; tail -c 400 gpurun_out/r02_c_bench_cfg$c.log; echo
done
