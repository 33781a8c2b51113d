#!/bin/bash
# r2af: device-side result sort + 256-slot reservations: GPU suite, default bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -12
timeout 400 python bench.py > gpurun_out/r2af_bench.json 2> gpurun_out/r2af_bench.err; echo "bench exit $?"; tail -3 gpurun_out/r2af_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2af_bench.json').read().strip().splitlines()[-1])
print('cfg2', round(d['value'],1), round(d['ms_per_step'],3), 'wall', round(d['value_wall'],1), round(d['wall_s_timed_region']/d['steps']*1000,3), 'e2e', round(d['e2e']['value'],2), d['host_us_last_step'], d['roofline']['kernel_ms_per_step'], 'frac', d['roofline']['frac'], d['roofline']['whole_path_frac'], d['parity']['counters_equal'], d['parity']['records_equal'], 'cpu', d['cpu_baseline']['value'], 'launches', d['gpu_launches'])
for k,v in d['per_config'].items(): print(k, round(v['value'],1), v['dominant_kernel'], round(v['dominant_kernel_frac'],3), v['parity']['records_equal'])
print('alt', d['alt_path']['value'])
PY
