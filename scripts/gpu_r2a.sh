# round 2, first GPU call: full-size parity (-m gpu incl. scale-1.0 tests), the 512-slot nondeterminism experiment, bench with the parity gate
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 ) > gpurun_out/r2a_pytest.log 2>&1; tail -25 gpurun_out/r2a_pytest.log
( time timeout 900 python scripts/gpu_nondet.py 8 5 ) > gpurun_out/r2a_nondet.log 2>&1; grep -E "distinct|oracle:" gpurun_out/r2a_nondet.log
( time timeout 900 python bench.py --steps 5 --warmup 3 ) > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; tail -c 1500 gpurun_out/r2a_bench.json; tail -3 gpurun_out/r2a_bench.err
