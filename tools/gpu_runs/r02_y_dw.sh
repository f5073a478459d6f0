#!/bin/bash
# round 2, call Y (1 GPU): dW GEMM with eight operand warps, re-measured after the shared-memory address-space fix
mkdir -p gpurun_out
for o in 4 8; do echo "== SFB200_TA_DW_OPW=$o"; SFB200_TA_DW_OPW=$o timeout 200 python tools/dw_bench.py 2>&1 | grep "^M=.*dW"; done | tee gpurun_out/r02_y_dw.log
