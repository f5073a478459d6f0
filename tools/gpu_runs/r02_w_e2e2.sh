#!/bin/bash
# round 2, call W (1 GPU): merged per-step graphs for host envs (D2H of the actions inside the graph), cheaper host simulator
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_engine.py tests/test_gpu_host_env.py tests/test_gpu_multi_policy.py tests/test_boundary.py -x -q -k "double_buffered or host_env or split_sampler or multi_agent or cartpole or gym_env" > gpurun_out/r02_w_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r02_w_pytest.log
for sp in 1 2; do timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-splits $sp > gpurun_out/r02_w_bench_s$sp.log 2>&1; echo "bench splits=$sp rc=$?"; grep -o '"e2e": {"value": [0-9.]*' gpurun_out/r02_w_bench_s$sp.log; grep -o '"async_rl": {"value": [0-9.]*' gpurun_out/r02_w_bench_s$sp.log | head -1; done
tail -5 gpurun_out/r02_w_bench_s1.log | cut -c1-600
