#!/bin/bash
# fused heads epilogue + post/pre fusion: new tests first, then everything, then bench
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "fused" > gpurun_out/fused_tests.log 2>&1; echo "fused tests rc=$?"; tail -25 gpurun_out/fused_tests.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/all_tests.log 2>&1; echo "all tests rc=$?"; tail -15 gpurun_out/all_tests.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench14.json 2> gpurun_out/bench14.err; echo "bench rc=$?"; tail -3 gpurun_out/bench14.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench14.json'))
print({k:d[k] for k in ['value','ms_per_step']}, d['e2e']['value'], d['async_rl']['value'], d['roofline']['achieved'], d['roofline_sampler']['rollout_ms'], d['launches_per_step'])
PY
