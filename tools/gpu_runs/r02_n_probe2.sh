#!/bin/bash
# round 2, call N (1 GPU): is the forward / dX GEMM paced by the L2 -> SM operand traffic?  (probe 16: no B_lo load; 32/64/128: L2 prefetch of A)
mkdir -p gpurun_out
for p in 0 16 32 64 128 20; do echo "== SFB200_TA_PROBE=$p"; SFB200_TA_PROBE=$p timeout 200 python tools/dw_bench.py 2>&1 | grep "^M=.*512 K=512"; done > gpurun_out/r02_n_probe.log 2>&1; cat gpurun_out/r02_n_probe.log
