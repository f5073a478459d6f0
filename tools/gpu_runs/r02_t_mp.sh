#!/bin/bash
# round 2, call T (1 GPU): multi-policy / PBT / multi-agent tests
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi_policy.py -x -q > gpurun_out/r02_t_pytest.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/r02_t_pytest.log
