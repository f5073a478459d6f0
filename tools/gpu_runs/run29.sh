#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/all_tests.log 2>&1; echo "all tests rc=$?"; tail -25 gpurun_out/all_tests.log | cut -c1-220
