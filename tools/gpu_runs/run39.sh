#!/bin/bash
# ncu --set full of the pipelined heads_backward kernel (+ the previous kernel for comparison)
mkdir -p gpurun_out
timeout 300 python tools/ncu_target4.py 2>&1 | tail -2
SFB200_HB_PIPE=0 timeout 300 python tools/ncu_target4.py 2>&1 | tail -1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"heads_backward_(pipe|vec)" -s 2 -c 1 -o gpurun_out/r01m_heads_backward_pipe python tools/ncu_target4.py > gpurun_out/ncu39.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu39.log
SFB200_HB_PIPE=0 timeout 600 ncu --set full --import-source on --clock-control none -k regex:"heads_backward_(pipe|vec)" -s 2 -c 1 -o gpurun_out/r01m_heads_backward_vec python tools/ncu_target4.py > gpurun_out/ncu39b.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/*.ncu-rep
