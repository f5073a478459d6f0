#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/all_tests.log 2>&1; echo "all tests rc=$?"; tail -4 gpurun_out/all_tests.log
timeout 300 python tools/gemm_bench.py > gpurun_out/gemm_bench.log 2>&1; grep -E "3xtf32" gpurun_out/gemm_bench.log
timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-async > gpurun_out/bench21.json 2> gpurun_out/bench21.err; echo "bench rc=$?"; tail -3 gpurun_out/bench21.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench21.json'))
print({k:d[k] for k in ['value','ms_per_step']}, d['roofline']['achieved'], d['roofline_sampler']['rollout_ms'], [(r['kernel'], round(r['avg_kernel_ms']*1e3,1)) for r in d['roofline_secondary']])
PY
