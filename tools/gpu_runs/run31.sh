#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_host_env.py -m gpu -q -x -k cartpole_learns > gpurun_out/cartpole.log 2>&1; echo "cartpole rc=$?"; tail -15 gpurun_out/cartpole.log | cut -c1-300
timeout 900 python bench.py > gpurun_out/bench31.json 2> gpurun_out/bench31.err; echo "bench rc=$?"; tail -2 gpurun_out/bench31.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench31.json'))
print({k:d[k] for k in ['value','ms_per_step']}, 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e'].get('async_rl'), 'async', d['async_rl']['value'], d['roofline']['achieved'], d['cpu_baseline']['value'])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r01k.csv python bench.py --steps 2 --warmup 3 --no-graph --no-e2e --no-cpu-baseline --no-async > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
