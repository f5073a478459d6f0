#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "heads_backward or heads" > gpurun_out/hb_tests.log 2>&1; echo "heads tests rc=$?"; tail -5 gpurun_out/hb_tests.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_ta|heads_backward_vec" -s 8 -c 4 -o gpurun_out/r01h_shortk python tools/ncu_target2.py > gpurun_out/ncu20.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu20.log
timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-async > gpurun_out/bench20.json 2> gpurun_out/bench20.err; echo "bench rc=$?"; tail -3 gpurun_out/bench20.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench20.json'))
print({k:d[k] for k in ['value','ms_per_step']}, d['roofline']['achieved'], d['roofline_sampler']['rollout_ms'], [(r['kernel'], round(r['avg_kernel_ms']*1e3,1)) for r in d['roofline_secondary']])
PY
