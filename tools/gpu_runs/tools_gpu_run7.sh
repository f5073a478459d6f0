#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sampling_api.py tests/test_gpu_kernels.py -m gpu -q -x -k "sampling or rnn" > gpurun_out/api_tests.log 2>&1; echo "api tests rc=$?"; tail -40 gpurun_out/api_tests.log
