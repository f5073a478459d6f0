#!/bin/bash
# heads_backward pipe kernel: up-front coefficient staging + L2 prefetch; variants
mkdir -p gpurun_out
for v in 0 1 2; do
  SFB200_HB_VARIANT=$v timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --timeout 600 -k "heads_backward" 2>&1 | tail -1
  SFB200_HB_VARIANT=$v timeout 300 python tools/ncu_target4.py 2>&1 | tail -1
done
timeout 600 python -m pytest tests/test_gpu_engine.py -m gpu -q -x --timeout 600 -k "learner_matches or graphed_learner or closed_loop" 2>&1 | tail -1
timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-async > gpurun_out/bench41.json 2> gpurun_out/bench41.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/bench41.json",):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    hb = [k for k in d.get("roofline_secondary", []) if k["kernel"] == "heads_backward"]
    print(f, d["value"], d["ms_per_step"], d["roofline"]["avg_kernel_ms"], hb[0]["avg_kernel_ms"] if hb else None, hb[0]["frac"] if hb else None)
PY
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"heads_backward_(pipe|vec)" -s 2 -c 1 -o gpurun_out/r01m_heads_backward_pipe3 python tools/ncu_target4.py > gpurun_out/ncu41.log 2>&1; echo "ncu rc=$?"
