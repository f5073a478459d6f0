#!/bin/bash
# pipelined heads_backward (A/B via SFB200_HB_PIPE), enjoy test fix, full GPU suite
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/all_tests.log 2>&1; echo "all tests rc=$?"
tail -6 gpurun_out/all_tests.log
timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-async > gpurun_out/bench38.json 2> gpurun_out/bench38.err; echo "bench rc=$?"
SFB200_HB_PIPE=0 timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-async > gpurun_out/bench38_nopipe.json 2> gpurun_out/bench38_nopipe.err; echo "bench nopipe rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/bench38.json", "gpurun_out/bench38_nopipe.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        hb = [k for k in d.get("roofline_secondary", []) if k["kernel"] == "heads_backward"]
        print(f, d["value"], d["ms_per_step"], d["roofline"]["avg_kernel_ms"], hb[0]["avg_kernel_ms"] if hb else None, hb[0]["frac"] if hb else None)
    except Exception as e:
        print(f, "ERR", e)
PY
