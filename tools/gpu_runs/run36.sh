#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/all_tests.log 2>&1; echo "all tests rc=$?"; tail -25 gpurun_out/all_tests.log | cut -c1-250
timeout 600 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/bench36.json 2> gpurun_out/bench36.err; echo "bench rc=$?"; tail -2 gpurun_out/bench36.err
SFB200_HEADS_FINISH_IN_GEMM=0 timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-async > gpurun_out/bench36_nofin.json 2> gpurun_out/bench36_nofin.err; echo "bench nofin rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/bench36.json','gpurun_out/bench36_nofin.json'):
    d=json.load(open(f))
    print(f, {k:d[k] for k in ['value','ms_per_step']}, d['async_rl'] and d['async_rl']['value'], d['roofline']['achieved'], d['roofline']['avg_kernel_ms'], d['roofline_sampler']['rollout_ms'], d['launches_per_step'])
PY
