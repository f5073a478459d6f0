#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "linear or tc_engine or rnn" > gpurun_out/gemm_tests.log 2>&1; echo "gemm tests rc=$?"; tail -12 gpurun_out/gemm_tests.log
timeout 300 python tools/gemm_bench.py 2>&1 | grep -E "3xtf32|tf32" | grep -v "^fwd.*simt"
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/all_tests.log 2>&1; echo "all tests rc=$?"; tail -4 gpurun_out/all_tests.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench12.json 2> gpurun_out/bench12.err; echo "bench rc=$?"; tail -3 gpurun_out/bench12.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench12.json'))
print({k:d[k] for k in ['value','ms_per_step']}, d['e2e']['value'], d['async_rl']['value'], d['roofline']['achieved'], d['roofline_sampler']['rollout_ms'])
PY
