#!/bin/bash
# 2-GPU session: data-parallel equivalence test + N=2 bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv
timeout 600 python -m pytest tests/test_gpu_multi.py -q -x -p no:cacheprovider > gpurun_out/pytest_multi.log 2>&1; echo "multi rc=$?"; tail -15 gpurun_out/pytest_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench25_n2.json 2> gpurun_out/bench25_n2.err; echo "bench n2 rc=$?"; tail -3 gpurun_out/bench25_n2.err | cut -c1-300; cut -c1-400 gpurun_out/bench25_n2.json
