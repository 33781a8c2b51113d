# the driver's round-end sequence on one GPU: smoke, full bench, reference arm, ncu launch list
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -c 2500 gpurun_out/bench_full.json; tail -3 gpurun_out/bench_full.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-300 gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_ref.err
B="python bench.py --gb 4 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
$B > gpurun_out/plain2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv $B > gpurun_out/ncu2.log 2>&1; tail -1 gpurun_out/ncu2.log | cut -c1-200
