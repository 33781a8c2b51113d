#!/bin/bash
# (historical: the parked / refill iptrie_kernel variants these runs compared were measured and then removed — profiles/README.md "last session"; MATCHY_B200_VARIANT=1 and MATCHY_B200_IPTRIE_MINB=3 select nothing in the committed library)
# r2ah: iptrie_kernel with parked walks (per-warp queue of unfinished walks) vs the r2f kernel (MATCHY_B200_VARIANT=1), configs 3 and 5
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/r2ah_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/r2ah_tests.log
for v in 0 1; do for c in 3 5; do
MATCHY_B200_VARIANT=$v timeout 300 python bench.py --config $c --gb 8 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-per-config --parity-gb 1 > gpurun_out/r2ah_c${c}_v$v.json 2> gpurun_out/r2ah_c${c}_v$v.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2ah_c${c}_v$v.json').read().strip().splitlines()[-1])
    print('variant $v cfg $c', round(d['value'],1), round(d['ms_per_step'],3), 'wall', round(d['value_wall'],1), {k:round(x,3) for k,x in d['roofline']['kernel_ms_per_step'].items()}, d['parity']['counters_equal'], d['parity']['records_equal'])
except Exception as e:
    print('variant $v cfg $c FAILED', e)
PY
done; done
