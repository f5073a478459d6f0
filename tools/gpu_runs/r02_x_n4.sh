#!/bin/bash
# round 2, call X' (4 GPUs): N = 4 bench line (peer-kernel exchanges among 4 ranks, dp_check, strong scaling) -- the 8-GPU slot was busy
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r02_x_bench_n4.log 2>&1
echo "bench n4 rc=$?"; grep -o '"value": [0-9.]*, "unit": "env-steps/s", "n_gpus": 4' gpurun_out/r02_x_bench_n4.log; grep -o '"dp_check": {[^}]*}' gpurun_out/r02_x_bench_n4.log | cut -c1-500; grep -o '"strong_scaling": {[^}]*}' gpurun_out/r02_x_bench_n4.log | cut -c1-200; grep -o '"e2e": {"value": [0-9.]*' gpurun_out/r02_x_bench_n4.log; tail -3 gpurun_out/r02_x_bench_n4.log | cut -c1-300
