#!/bin/bash
# round 2, call U (1 GPU): final build -- smoke(), full GPU suite, full bench line, ncu launch list + two --set full captures
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/r02_u_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02_u_smoke.log
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r02_u_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_u_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_u_bench_n1.log 2>&1; echo "bench rc=$?"
grep -o '"value": [0-9.]*, "unit": "env-steps/s", "n_gpus": 1' gpurun_out/r02_u_bench_n1.log; grep -o '"e2e": {"value": [0-9.]*' gpurun_out/r02_u_bench_n1.log; grep -o '"cpu_baseline": {"value": [0-9.]*' gpurun_out/r02_u_bench_n1.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 250 -c 600 --csv --log-file gpurun_out/r02_u_launches.csv python bench.py --steps 2 --warmup 3 --no-graph --no-e2e --no-cpu-baseline --no-async > gpurun_out/r02_u_ncu_list.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/r02_u_launches.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"rollout_mlp2_tape_kernel|gemm_tc_ta_kernel" -s 40 -c 8 -o gpurun_out/r02_u_full python bench.py --steps 2 --warmup 3 --no-graph --no-e2e --no-cpu-baseline --no-async > gpurun_out/r02_u_ncu_full.log 2>&1
echo "ncu full rc=$?"; ls -la gpurun_out/r02_u_full.ncu-rep
