#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_engine.py tests/test_gpu_host_env.py -m gpu -q -x -k "split or host_env" > gpurun_out/new_tests.log 2>&1; echo "new tests rc=$?"; tail -25 gpurun_out/new_tests.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/all_tests.log 2>&1; echo "all tests rc=$?"; tail -6 gpurun_out/all_tests.log
for sp in 2 1 4; do
timeout 600 python bench.py --no-cpu-baseline --no-e2e --splits $sp > gpurun_out/bench24_s$sp.json 2> gpurun_out/bench24_s$sp.err; echo "bench splits=$sp rc=$?"; tail -2 gpurun_out/bench24_s$sp.err
done
python - <<'PY'
import json
for sp in (2,1,4):
    try:
        d=json.load(open(f'gpurun_out/bench24_s{sp}.json'))
        print(sp, {k:d[k] for k in ['value','ms_per_step']}, d['async_rl'] and d['async_rl']['value'], d['roofline']['achieved'], d['roofline_sampler']['rollout_ms'], d['launches_per_step'])
    except Exception as e: print(sp, 'ERR', e)
PY
