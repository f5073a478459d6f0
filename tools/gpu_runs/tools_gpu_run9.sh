#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "linear or tc_engine or rnn" > gpurun_out/gemm_tests.log 2>&1; echo "gemm tests rc=$?"; tail -30 gpurun_out/gemm_tests.log
timeout 300 python tools/gemm_bench.py > gpurun_out/gemm_bench_ta.log 2>&1; echo "gemm bench rc=$?"; cat gpurun_out/gemm_bench_ta.log | tail -30
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/all_tests.log 2>&1; echo "all tests rc=$?"; tail -8 gpurun_out/all_tests.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench9.json 2> gpurun_out/bench9.err; echo "bench rc=$?"; tail -3 gpurun_out/bench9.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench9.json'))
print({k:d[k] for k in ['value','ms_per_step']}, d['e2e']['value'], d['async_rl']['value'], d['roofline']['achieved'], d['roofline_sampler']['rollout_ms'])
PY
