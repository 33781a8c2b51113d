mkdir -p gpurun_out
TAG=${1:-r2r}
export MATCHY_B200_WIDE_SCAN=1 MATCHY_B200_WIDE_WARPS=28
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest.log 2>&1; tail -4 gpurun_out/${TAG}_pytest.log
for c in 2 3 4 1 5; do
timeout 600 python bench.py --config $c --gb 8 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_c$c.json 2> gpurun_out/${TAG}_c$c.err; python -c "
import json
d=json.load(open('gpurun_out/${TAG}_c$c.json')); print('wide28', $c, round(d['value'],1), round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['roofline']['kernel_ms_per_step'].items()}, d['counters']['matches'], d['parity']['counters_equal'], d['parity']['records_equal'], round(d['wall_s_timed_region']*1000/3,2))"; tail -2 gpurun_out/${TAG}_c$c.err; done
unset MATCHY_B200_WIDE_SCAN MATCHY_B200_WIDE_WARPS
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_default10.json 2> gpurun_out/${TAG}_default10.err; python -c "
import json
d=json.load(open('gpurun_out/${TAG}_default10.json')); print('default fused hot16 10GB', round(d['value'],1), round(d['ms_per_step'],3))"
MATCHY_B200_WIDE_SCAN=1 MATCHY_B200_WIDE_WARPS=28 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_wide10.json 2> gpurun_out/${TAG}_wide10.err; python -c "
import json
d=json.load(open('gpurun_out/${TAG}_wide10.json')); print('wide28 10GB', round(d['value'],1), round(d['ms_per_step'],3))"
MATCHY_B200_UNFUSED=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_unfused10.json 2> gpurun_out/${TAG}_unfused10.err; python -c "
import json
d=json.load(open('gpurun_out/${TAG}_unfused10.json')); print('unfused 10GB', round(d['value'],1), round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['roofline']['kernel_ms_per_step'].items()})"
