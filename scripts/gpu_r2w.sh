#!/bin/bash
# r2w: full GPU suite with the new mode tests, then the default bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2w_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/r2w_tests.log
tail -5 gpurun_out/r2w_tests.log
timeout 400 python bench.py > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err; echo "bench exit $?"
cat gpurun_out/r2w_bench.json
