#!/bin/bash
# round 2, call A (2 GPUs): NVLink peer-memory exchanges -- micro-benchmark, equivalence tests, N=2 bench
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_a_topo.txt 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_comm_bench.py > gpurun_out/r02_a_dp_comm_bench.log 2>&1
echo "dp_comm_bench rc=$?"
grep DP_COMM_BENCH gpurun_out/r02_a_dp_comm_bench.log | tail -1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_a_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r02_a_pytest_gpu.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_a_bench_n2.log 2>&1
echo "bench n2 rc=$?"; tail -c 1500 gpurun_out/r02_a_bench_n2.log
