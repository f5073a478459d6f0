#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sampling_api.py tests/test_gpu_engine.py tests/test_gpu_kernels.py -m gpu -q -k "sampling or rnn or async" > gpurun_out/api_tests.log 2>&1; echo "api tests rc=$?"; tail -40 gpurun_out/api_tests.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench8.json 2> gpurun_out/bench8.err; echo "bench rc=$?"; tail -5 gpurun_out/bench8.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench8.json'))
print({k:d[k] for k in ['value','ms_per_step']}, d['e2e']['value'], d['async_rl'], d['roofline_sampler'])
PY
