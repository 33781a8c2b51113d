mkdir -p gpurun_out
TAG=${1:-r2j}
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest.log 2>&1; tail -4 gpurun_out/${TAG}_pytest.log
for c in 2 1 3 4 5; do timeout 600 python bench.py --config $c --gb 8 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_bench_c$c.json 2> gpurun_out/${TAG}_bench_c$c.err; python -c "
import json
d=json.load(open('gpurun_out/${TAG}_bench_c$c.json')); print($c, round(d['value'],1), round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['roofline']['kernel_ms_per_step'].items()}, d['counters']['matches'], d['parity']['counters_equal'], d['parity']['records_equal'])"; tail -2 gpurun_out/${TAG}_bench_c$c.err; done
B="python bench.py --gb 0.5 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-parity --chunk-mb 512"
$B > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'scan_kernel' -c 1 -f -o gpurun_out/prof_$TAG $B > gpurun_out/ncu.log 2>&1; tail -2 gpurun_out/ncu.log
