#!/bin/bash
# final validation of the round-1 build: full GPU suite, smoke, full bench (cpu baseline + e2e + async), reference arm
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/all_tests.log 2>&1; echo "all tests rc=$?"; tail -2 gpurun_out/all_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench42.json 2> gpurun_out/bench42.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench42.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench42_ref.json 2> gpurun_out/bench42_ref.err; echo "ref rc=$?"; tail -c 400 gpurun_out/bench42_ref.json
