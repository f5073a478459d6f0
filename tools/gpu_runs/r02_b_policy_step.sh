#!/bin/bash
# round 2, call B (1 GPU): fused two-layer policy step -- kernel test first (bounded), then the full suite and the bench
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -k "policy_mlp2" > gpurun_out/r02_b_pytest_ps.log 2>&1
echo "policy_mlp2 tests rc=$?"; tail -15 gpurun_out/r02_b_pytest_ps.log
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02_b_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -25 gpurun_out/r02_b_pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_b_bench_n1.log 2>&1
echo "bench rc=$?"; tail -c 600 gpurun_out/r02_b_bench_n1.log
SFB200_POLICY_FUSED=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-async > gpurun_out/r02_b_bench_n1_unfused.log 2>&1
echo "bench unfused rc=$?"
