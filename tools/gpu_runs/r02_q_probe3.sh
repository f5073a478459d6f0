#!/bin/bash
# round 2, call Q (1 GPU): how much of a GEMM tile is the epilogue's arithmetic + stores (probe 256 drops them)
mkdir -p gpurun_out
for p in 0 256; do echo "== SFB200_TA_PROBE=$p"; SFB200_TA_PROBE=$p timeout 200 python tools/dw_bench.py 2>&1 | grep "^M=.*\(fwd\|fp16\)"; done | tee gpurun_out/r02_q_probe.log
