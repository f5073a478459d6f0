#!/bin/bash
# 2-GPU session: data-parallel equivalence test + scaling bench (N=1 and N=2 back to back)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv
timeout 600 python -m pytest tests/test_gpu_multi.py -q -x --timeout 500 -p no:cacheprovider > gpurun_out/pytest_multi.log 2>&1; echo "multi rc=$?"; tail -15 gpurun_out/pytest_multi.log
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.log 2>&1; echo "bench n1 rc=$?"; tail -1 gpurun_out/bench_n1.log | cut -c1-200
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2.log 2>&1; echo "bench n2 rc=$?"; tail -3 gpurun_out/bench_n2.log | cut -c1-400
