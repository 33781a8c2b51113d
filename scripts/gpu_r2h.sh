mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2h_pytest.log 2>&1; tail -15 gpurun_out/r2h_pytest.log
for c in 2 1 3 4 5; do timeout 600 python bench.py --config $c --gb 8 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2h_bench_c$c.json 2> gpurun_out/r2h_bench_c$c.err; python -c "
import json
d=json.load(open('gpurun_out/r2h_bench_c$c.json')); print($c, round(d['value'],1), round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['roofline']['kernel_ms_per_step'].items()}, d['counters'], d['parity']['counters_equal'], d['parity']['records_equal'])"; tail -2 gpurun_out/r2h_bench_c$c.err; done
