#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/all_tests.log 2>&1; echo "all tests rc=$?"; tail -12 gpurun_out/all_tests.log | cut -c1-220
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench30.json 2> gpurun_out/bench30.err; echo "bench rc=$?"; tail -2 gpurun_out/bench30.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench30.json'))
print({k:d[k] for k in ['value','ms_per_step']}, 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'async', d['async_rl']['value'], d['roofline']['achieved'], d['roofline_sampler']['rollout_ms'], d['launches_per_step'])
PY
