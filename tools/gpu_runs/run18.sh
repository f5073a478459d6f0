#!/bin/bash
# continuous (Gaussian) action path: kernel tests, golden engine tests, then the whole suite
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "continuous" > gpurun_out/cont_tests.log 2>&1; echo "continuous kernel tests rc=$?"; tail -30 gpurun_out/cont_tests.log
timeout 600 python -m pytest tests/test_gpu_engine.py -m gpu -q -k "gauss" > gpurun_out/gauss_engine.log 2>&1; echo "gauss engine tests rc=$?"; tail -40 gpurun_out/gauss_engine.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/all_tests.log 2>&1; echo "all tests rc=$?"; tail -8 gpurun_out/all_tests.log
