#!/bin/bash
# 2-GPU session with the final build: whole GPU suite (incl. the NCCL data-parallel equivalence test), N=2 and N=1 bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/all_tests.log 2>&1; echo "all tests rc=$?"; tail -6 gpurun_out/all_tests.log | cut -c1-200
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench32_n2.json 2> gpurun_out/bench32_n2.err; echo "bench n2 rc=$?"; tail -2 gpurun_out/bench32_n2.err | cut -c1-200; cut -c1-330 gpurun_out/bench32_n2.json
timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-async > gpurun_out/bench32_n1.json 2> gpurun_out/bench32_n1.err; echo "bench n1 rc=$?"; cut -c1-330 gpurun_out/bench32_n1.json
