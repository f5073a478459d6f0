#!/bin/bash
# round 2, call S (2 GPUs): the 2-GPU equivalence tests + N=2 bench (dp_check, strong scaling) + N=1 bench with the e2e arm, fp16-split engine
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/r02_s_pytest_multi.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r02_s_pytest_multi.log
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_s_bench_n2.log 2>&1
echo "bench n2 rc=$?"; grep -o '"value": [0-9.]*, "unit": "env-steps/s", "n_gpus": 2' gpurun_out/r02_s_bench_n2.log; grep -o '"dp_check": {[^}]*}' gpurun_out/r02_s_bench_n2.log; grep -o '"strong_scaling": {[^}]*}' gpurun_out/r02_s_bench_n2.log | cut -c1-300
CUDA_VISIBLE_DEVICES=0 timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_s_bench_n1.log 2>&1
echo "bench n1 rc=$?"; grep -o '"value": [0-9.]*, "unit": "env-steps/s", "n_gpus": 1' gpurun_out/r02_s_bench_n1.log; grep -o '"e2e": {"value": [0-9.]*' gpurun_out/r02_s_bench_n1.log; grep -o '"async_rl": {"value": [0-9.]*' gpurun_out/r02_s_bench_n1.log
