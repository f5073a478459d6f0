#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "rnn or gru or lstm" > gpurun_out/rnn_tests.log 2>&1; echo "rnn tests rc=$?"; tail -30 gpurun_out/rnn_tests.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/all_tests.log 2>&1; echo "all tests rc=$?"; tail -15 gpurun_out/all_tests.log
timeout 600 python bench.py > gpurun_out/bench6.json 2> gpurun_out/bench6.err; echo "bench rc=$?"; cat gpurun_out/bench6.json
