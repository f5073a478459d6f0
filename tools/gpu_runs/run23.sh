#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/all_tests.log 2>&1; echo "all tests rc=$?"; tail -8 gpurun_out/all_tests.log
timeout 900 python bench.py > gpurun_out/bench23.json 2> gpurun_out/bench23.err; echo "bench rc=$?"; tail -3 gpurun_out/bench23.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench23.json'))
print({k:d[k] for k in ['value','ms_per_step']}, d['e2e']['value'], d['async_rl']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline_sampler']['rollout_ms'], d['cpu_baseline']['value'], d['launches_per_step'])
PY
