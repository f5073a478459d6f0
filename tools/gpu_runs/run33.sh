#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none -k regex:"normalize_kernel|gae_returns|post_pre_step|ppo_loss_kernel|clip_adam|heads_from_partials|moments_partial" -s 7 -c 7 -o gpurun_out/r01k_elementwise python tools/ncu_target3.py > gpurun_out/ncu33.log 2>&1; echo "ncu rc=$?"; tail -4 gpurun_out/ncu33.log
ncu -i gpurun_out/r01k_elementwise.ncu-rep --page raw --csv > gpurun_out/r01k_elementwise_raw.csv 2>/dev/null; wc -l gpurun_out/r01k_elementwise_raw.csv
