#!/bin/bash
# round 2, call V (1 GPU): double-buffered host sampling (test + e2e arm), multi-block bound kernels, ncu --set full of the rollout kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_engine.py tests/test_gpu_kernels.py tests/test_gpu_host_env.py -x -q -k "double_buffered or out_bound or fp16_split or host_env or split_sampler or graphed_learner" > gpurun_out/r02_v_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r02_v_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_v_bench.log 2>&1; echo "bench rc=$?"
grep -o '"value": [0-9.]*, "unit": "env-steps/s", "n_gpus": 1' gpurun_out/r02_v_bench.log; grep -o '"e2e": {"value": [0-9.]*' gpurun_out/r02_v_bench.log; grep -o '"async_rl": {"value": [0-9.]*' gpurun_out/r02_v_bench.log; tail -3 gpurun_out/r02_v_bench.log | cut -c1-400
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-splits 1 --no-async > gpurun_out/r02_v_bench_s1.log 2>&1; grep -o '"e2e": {"value": [0-9.]*' gpurun_out/r02_v_bench_s1.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"rollout_mlp2_tape_kernel" -s 2 -c 1 -o gpurun_out/r02_v_rollout python bench.py --steps 2 --warmup 3 --no-graph --no-e2e --no-cpu-baseline --no-async > gpurun_out/r02_v_ncu.log 2>&1; echo "ncu rc=$?"; ls -la gpurun_out/r02_v_rollout.ncu-rep
