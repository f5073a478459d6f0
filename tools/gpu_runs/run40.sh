#!/bin/bash
# FFMA2 / low-instruction pipelined heads_backward
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_engine.py -m gpu -q -x --timeout 600 -k "heads_backward or learner_matches or graphed_learner or closed_loop" > gpurun_out/hb_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/hb_tests.log
timeout 300 python tools/ncu_target4.py 2>&1 | tail -1
timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-async > gpurun_out/bench40.json 2> gpurun_out/bench40.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/bench40.json",):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    hb = [k for k in d.get("roofline_secondary", []) if k["kernel"] == "heads_backward"]
    print(f, d["value"], d["ms_per_step"], d["roofline"]["avg_kernel_ms"], hb[0]["avg_kernel_ms"] if hb else None, hb[0]["frac"] if hb else None)
PY
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"heads_backward_(pipe|vec)" -s 2 -c 1 -o gpurun_out/r01m_heads_backward_pipe2 python tools/ncu_target4.py > gpurun_out/ncu40.log 2>&1; echo "ncu rc=$?"
