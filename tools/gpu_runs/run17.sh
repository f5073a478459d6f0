#!/bin/bash
# pre-split weight lo (BLO): tests, gemm bench, bench with / without BLO
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/all_tests.log 2>&1; echo "all tests rc=$?"; tail -6 gpurun_out/all_tests.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench17.json 2> gpurun_out/bench17.err; echo "bench rc=$?"; tail -3 gpurun_out/bench17.err
SFB200_TC_B_LO=0 timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-async > gpurun_out/bench17_noblo.json 2> gpurun_out/bench17_noblo.err; echo "bench noblo rc=$?"; tail -3 gpurun_out/bench17_noblo.err
python - <<'PY'
import json
for f in ['gpurun_out/bench17.json','gpurun_out/bench17_noblo.json']:
    d=json.load(open(f))
    print(f, {k:d[k] for k in ['value','ms_per_step']}, d['e2e'] and d['e2e']['value'], d['async_rl'] and d['async_rl']['value'], d['roofline']['achieved'], d['roofline']['avg_kernel_ms'], d['roofline_sampler']['rollout_ms'], d['launches_per_step'])
PY
