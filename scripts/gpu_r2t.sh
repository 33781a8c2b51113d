mkdir -p gpurun_out
TAG=${1:-r2t}
( timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest.log 2>&1; tail -4 gpurun_out/${TAG}_pytest.log
for c in 2 3 4 1 5; do
timeout 600 python bench.py --config $c --gb 8 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_c$c.json 2> gpurun_out/${TAG}_c$c.err; python -c "
import json
d=json.load(open('gpurun_out/${TAG}_c$c.json')); print('default', $c, round(d['value'],1), round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['roofline']['kernel_ms_per_step'].items()}, d['counters']['matches'], d['parity']['counters_equal'], d['parity']['records_equal'], round(d['wall_s_timed_region']*1000/3,2))"; tail -2 gpurun_out/${TAG}_c$c.err; done
