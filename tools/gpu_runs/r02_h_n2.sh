#!/bin/bash
# round 2, call H (2 GPUs): full -m gpu suite (incl. the 2-GPU equivalence tests), dp micro-benchmark, N=2 bench with dp_check +
# strong scaling, N=1 bench (no CPU arm) for the scaling efficiency on the same box
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_h_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/r02_h_pytest_gpu.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_comm_bench.py > gpurun_out/r02_h_dp_comm_bench.log 2>&1
echo "dp_comm_bench rc=$?"; grep DP_COMM_BENCH gpurun_out/r02_h_dp_comm_bench.log | tail -1 | cut -c1-700
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_h_bench_n2.log 2>&1
echo "bench n2 rc=$?"; grep -o '"value": [0-9.]*, "unit": "env-steps/s", "n_gpus": 2' gpurun_out/r02_h_bench_n2.log; grep -o '"dp_check": {[^}]*}' gpurun_out/r02_h_bench_n2.log; grep -o '"strong_scaling": {[^}]*}' gpurun_out/r02_h_bench_n2.log | cut -c1-300
CUDA_VISIBLE_DEVICES=0 timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_h_bench_n1.log 2>&1
echo "bench n1 rc=$?"; grep -o '"value": [0-9.]*, "unit": "env-steps/s", "n_gpus": 1' gpurun_out/r02_h_bench_n1.log; grep -o '"e2e": {"value": [0-9.]*' gpurun_out/r02_h_bench_n1.log
