#!/bin/bash
# r2aa: iptrie register budgets on config 3 (and 5), host time breakdown of a config 2 step
mkdir -p gpurun_out
for mb in 4 5 6; do for c in 3 5; do
MATCHY_B200_IPTRIE_MINB=$mb timeout 300 python bench.py --config $c --gb 8 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-per-config --parity-gb 1 > gpurun_out/r2aa_c${c}_mb$mb.json 2> gpurun_out/r2aa_c${c}_mb$mb.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2aa_c${c}_mb$mb.json').read().strip().splitlines()[-1])
print('minb $mb cfg $c', round(d['value'],1), round(d['ms_per_step'],3), 'wall', round(d['value_wall'],1), {k:round(v,3) for k,v in d['roofline']['kernel_ms_per_step'].items()}, d['host_us_last_step'], d['parity']['counters_equal'], d['parity']['records_equal'])
PY
done; done
timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-per-config --parity-gb 1 > gpurun_out/r2aa_c2.json 2> gpurun_out/r2aa_c2.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2aa_c2.json').read().strip().splitlines()[-1])
print('cfg 2', round(d['value'],1), round(d['ms_per_step'],3), 'wall', round(d['value_wall'],1), d['wall_s_timed_region']/d['steps']*1000, d['host_us_last_step'])
PY
