#!/bin/bash
# round 2, call P (1 GPU): the fp16-split GEMM engine (kernel-level accuracy + timing), 4 vs 8 operand warps
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -s -k "fp16_split or out_bound" > gpurun_out/r02_p_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|Error|error" gpurun_out/r02_p_pytest.log | head -30
for o in 8 4; do echo "== SFB200_F16_OPW=$o"; SFB200_F16_OPW=$o timeout 200 python tools/dw_bench.py 2>&1 | grep "^M=.*fp16"; done | tee gpurun_out/r02_p_bench.log
