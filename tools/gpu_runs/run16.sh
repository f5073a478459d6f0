#!/bin/bash
# ncu --set full of the learner's layer-2 GEMMs (fused fwd, dW, dx), with source
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_ta -s 6 -c 3 -o gpurun_out/r01g_gemm_ta python tools/ncu_target.py > gpurun_out/ncu16.log 2>&1; echo "ncu rc=$?"; tail -5 gpurun_out/ncu16.log; ls -la gpurun_out/*.ncu-rep
