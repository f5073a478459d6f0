#!/bin/bash
# round 2, call X (8 GPUs): the bench line the driver's scaling run asks for, N = 8 (peer-kernel exchanges among 8 ranks, dp_check, strong scaling)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_x_bench_n8.log 2>&1
echo "bench n8 rc=$?"; grep -o '"value": [0-9.]*, "unit": "env-steps/s", "n_gpus": 8' gpurun_out/r02_x_bench_n8.log; grep -o '"dp_check": {[^}]*}' gpurun_out/r02_x_bench_n8.log | cut -c1-500; grep -o '"strong_scaling": {[^}]*}' gpurun_out/r02_x_bench_n8.log | cut -c1-200; grep -o '"e2e": {"value": [0-9.]*' gpurun_out/r02_x_bench_n8.log; tail -3 gpurun_out/r02_x_bench_n8.log | cut -c1-300
