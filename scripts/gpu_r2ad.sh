#!/bin/bash
# r2ad: sub-block interleave of the dotted / numeric passes in token_kernel: parity suite, then configs 2, 1, 5, 3 at several sub-block sizes
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for sb in 0 8192 16384 32768; do for c in 2 5 1; do
MATCHY_B200_SUB_BLOCK=$sb timeout 300 python bench.py --config $c --gb 8 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-per-config --parity-gb 1 > gpurun_out/r2ad_c${c}_sb$sb.json 2> gpurun_out/r2ad_c${c}_sb$sb.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2ad_c${c}_sb$sb.json').read().strip().splitlines()[-1])
print('sub_block $sb cfg $c', round(d['value'],1), round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['roofline']['kernel_ms_per_step'].items()}, d['parity']['counters_equal'], d['parity']['records_equal'])
PY
done; done
