#!/bin/bash
# r2ab: host-side result path after the pool changes (config 2 default bench, config 3 and 5 at 8 GB), GPU suite
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 400 python bench.py > gpurun_out/r2ab_bench.json 2> gpurun_out/r2ab_bench.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2ab_bench.json').read().strip().splitlines()[-1])
print('cfg2', round(d['value'],1), round(d['ms_per_step'],3), 'wall', round(d['value_wall'],1), round(d['wall_s_timed_region']/d['steps']*1000,3), 'e2e', round(d['e2e']['value'],2), d['host_us_last_step'], d['roofline']['kernel_ms_per_step'], d['parity']['counters_equal'], d['parity']['records_equal'])
for k,v in d['per_config'].items(): print(k, round(v['value'],1), v['parity']['records_equal'])
print('alt', d['alt_path']['value'])
PY
for c in 3 5; do
timeout 300 python bench.py --config $c --gb 8 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-per-config --parity-gb 1 > gpurun_out/r2ab_c$c.json 2> gpurun_out/r2ab_c$c.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2ab_c$c.json').read().strip().splitlines()[-1])
print('cfg $c', round(d['value'],1), round(d['ms_per_step'],3), 'wall', round(d['value_wall'],1), {k:round(v,3) for k,v in d['roofline']['kernel_ms_per_step'].items()}, d['host_us_last_step'], d['parity']['counters_equal'], d['parity']['records_equal'])
PY
done
