#!/bin/bash
# 2-GPU session: data-parallel equivalence test + N=2 bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -q -x --timeout 500 -p no:cacheprovider > gpurun_out/pytest_multi.log 2>&1; echo "multi rc=$?"; tail -5 gpurun_out/pytest_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench44_n2.json 2> gpurun_out/bench44_n2.err; echo "bench n2 rc=$?"; tail -c 500 gpurun_out/bench44_n2.json
