#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_configs.py -m gpu -q > gpurun_out/cfg_tests.log 2>&1; echo "config tests rc=$?"; tail -40 gpurun_out/cfg_tests.log | cut -c1-250
