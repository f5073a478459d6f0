#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "conv or uint8" > gpurun_out/conv_tests.log 2>&1; echo "conv kernel tests rc=$?"; tail -30 gpurun_out/conv_tests.log
timeout 600 python -m pytest tests/test_gpu_engine.py -m gpu -q -k "conv" > gpurun_out/conv_engine.log 2>&1; echo "conv engine tests rc=$?"; tail -40 gpurun_out/conv_engine.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/all_tests.log 2>&1; echo "all tests rc=$?"; tail -8 gpurun_out/all_tests.log
