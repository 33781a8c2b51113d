#!/bin/bash
# r2ae: ncu captures of the final code (reports reduced to CSV on the box), token reservation unit on configs 3 and 5
mkdir -p gpurun_out
bash scripts/gpu_r2_prof.sh r2f
for tu in 128 256 512; do for c in 3 5; do
MATCHY_B200_TOK_RESERVE=$tu timeout 300 python bench.py --config $c --gb 8 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-per-config --parity-gb 1 > gpurun_out/r2ae_c${c}_tu$tu.json 2> gpurun_out/r2ae_c${c}_tu$tu.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2ae_c${c}_tu$tu.json').read().strip().splitlines()[-1])
print('tok_reserve $tu cfg $c', round(d['value'],1), round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['roofline']['kernel_ms_per_step'].items()}, d['parity']['counters_equal'], d['parity']['records_equal'])
PY
done; done
du -sh gpurun_out
