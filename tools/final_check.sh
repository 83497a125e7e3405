#!/bin/bash
# Runs on the GPU box: what the driver runs at round end (default bench line, GPU tests, smoke, reference arm), each timed.
mkdir -p gpurun_out
( time timeout 400 python bench.py ) > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
echo "bench rc=$?"; cut -c1-300 gpurun_out/final_bench.json; tail -4 gpurun_out/final_bench.err
( time timeout 400 python -m pytest tests -m gpu -x -q ) > gpurun_out/final_pytest.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/final_pytest.log
( time timeout 120 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/final_smoke.log 2>&1
echo "smoke rc=$?"; tail -5 gpurun_out/final_smoke.log
( time timeout 200 python bench.py --impl reference ) > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err
echo "ref rc=$?"; cut -c1-200 gpurun_out/final_ref.json; tail -4 gpurun_out/final_ref.err
