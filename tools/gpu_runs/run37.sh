#!/bin/bash
# action masks + deterministic eval + enjoy; policy-lag snapshot fix (cfg3)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/all_tests.log 2>&1; echo "all tests rc=$?"
tail -25 gpurun_out/all_tests.log
