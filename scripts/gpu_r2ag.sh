#!/bin/bash
# r2ag: what slowed config 4 down: the reservation unit or the sort kernel?
mkdir -p gpurun_out
for tu in 128 256; do for hs in 0 1; do for c in 4 1; do
if [ $hs = 1 ]; then export MATCHY_B200_HOST_SORT=1; else unset MATCHY_B200_HOST_SORT; fi
MATCHY_B200_TOK_RESERVE=$tu timeout 300 python bench.py --config $c --gb 4 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-per-config --parity-gb 1 > gpurun_out/r2ag.json 2> gpurun_out/r2ag.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2ag.json').read().strip().splitlines()[-1])
print('tok_reserve $tu host_sort $hs cfg $c', round(d['value'],1), round(d['ms_per_step'],3), 'wall', round(d['value_wall'],1), {k:round(v,3) for k,v in d['roofline']['kernel_ms_per_step'].items()}, d['host_us_last_step'], d['parity']['records_equal'])
PY
done; done; done
