#!/bin/bash
# round 2, call F (1 GPU): ncu launch lists of the bench step (tail fused / separate; policy fused), cfg4 parity re-run
mkdir -p gpurun_out
for mode in "0 1" "0 0" "1 1"; do
  set -- $mode
  SFB200_POLICY_FUSED=$1 SFB200_TAIL_FUSED=$2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 1200 --csv --log-file gpurun_out/r02_f_launches_pf$1_tf$2.csv python bench.py --steps 2 --warmup 3 --no-graph --no-e2e --no-cpu-baseline --no-async > gpurun_out/r02_f_ncu_pf$1_tf$2.log 2>&1
  echo "launch list pf=$1 tf=$2 rc=$?"; wc -l gpurun_out/r02_f_launches_pf$1_tf$2.csv
done
timeout 300 python -m pytest tests/test_gpu_fullsize.py -q -k cfg4 > gpurun_out/r02_f_pytest_cfg4.log 2>&1; echo "cfg4 rc=$?"; tail -4 gpurun_out/r02_f_pytest_cfg4.log
