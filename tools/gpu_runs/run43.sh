#!/bin/bash
# ncu launch list of the final round-1 build (same command as the bench, eager so every launch is visible)
mkdir -p gpurun_out
timeout 700 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r01m.csv python bench.py --steps 2 --warmup 3 --no-graph --no-e2e --no-cpu-baseline --no-async > gpurun_out/ncu_bench43.log 2>&1; echo "ncu rc=$?"
python profiles/summarize_launches.py gpurun_out/launches_r01m.csv > gpurun_out/launches_r01m.md; head -30 gpurun_out/launches_r01m.md
# A/B: learner replayed as one CUDA graph in the device-resident arm
timeout 400 python bench.py --no-cpu-baseline --no-e2e --no-async --learner-graph > gpurun_out/bench43_lgraph.json 2> gpurun_out/bench43_lgraph.err; echo "lgraph rc=$?"
timeout 400 python bench.py --no-cpu-baseline --no-e2e --no-async > gpurun_out/bench43_eager.json 2> gpurun_out/bench43_eager.err; echo "eager rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/bench43_lgraph.json", "gpurun_out/bench43_eager.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["roofline_sampler"]["rollout_ms"] if d.get("roofline_sampler") else None)
    except Exception as e:
        print(f, "ERR", e)
PY
