#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/all_tests.log 2>&1; echo "all tests rc=$?"; tail -6 gpurun_out/all_tests.log
timeout 300 python tools/gemm_bench.py > gpurun_out/gemm_bench.log 2>&1; grep -E "32768.*3xtf32" gpurun_out/gemm_bench.log
SFB200_TC_XB=0 timeout 300 python tools/gemm_bench.py > gpurun_out/gemm_bench_noxb.log 2>&1; echo "no XB:"; grep -E "32768.*3xtf32" gpurun_out/gemm_bench_noxb.log
timeout 600 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/bench27.json 2> gpurun_out/bench27.err; echo "bench rc=$?"; tail -2 gpurun_out/bench27.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench27.json'))
print({k:d[k] for k in ['value','ms_per_step']}, d['async_rl'] and d['async_rl']['value'], d['roofline']['achieved'], d['roofline_sampler']['rollout_ms'], d['launches_per_step'])
PY
