#!/bin/bash
# 2-GPU session: data-parallel equivalence (eager + graph-captured learner) + N=2 bench with and without the DP learner graph
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_multi.py -q -x --timeout 400 -p no:cacheprovider > gpurun_out/pytest_multi.log 2>&1; echo "multi rc=$?"; tail -12 gpurun_out/pytest_multi.log | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench44_n2.json 2> gpurun_out/bench44_n2.err; echo "bench n2 (dp graph) rc=$?"; tail -c 300 gpurun_out/bench44_n2.err
SFB200_DP_GRAPH=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-async > gpurun_out/bench44_n2_eager.json 2> gpurun_out/bench44_n2_eager.err; echo "bench n2 (eager learner) rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/bench44_n2.json", "gpurun_out/bench44_n2_eager.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["n_gpus"], (d.get("e2e") or {}).get("value"), (d.get("async_rl") or {}).get("value"))
    except Exception as e:
        print(f, "ERR", e)
PY
