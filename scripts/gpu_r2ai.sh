#!/bin/bash
# (historical: the parked / refill iptrie_kernel variants these runs compared were measured and then removed — profiles/README.md "last session"; MATCHY_B200_VARIANT=1 and MATCHY_B200_IPTRIE_MINB=3 select nothing in the committed library)
# r2ai: iptrie_kernel with persistent per-lane walks refilled from a per-warp queue of parked walks: 64 registers (spills) vs 80
# registers (3 blocks per SM) vs the r2f kernel (MATCHY_B200_VARIANT=1), configs 3 and 5
mkdir -p gpurun_out
run() {  # name, config, env...
  local name=$1 c=$2; shift 2
  env "$@" timeout 300 python bench.py --config $c --gb 8 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-per-config --parity-gb 1 > gpurun_out/r2ai_c${c}_$name.json 2> gpurun_out/r2ai_c${c}_$name.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2ai_c${c}_$name.json').read().strip().splitlines()[-1])
    print('$name cfg $c', round(d['value'],1), round(d['ms_per_step'],3), 'wall', round(d['value_wall'],1), {k:round(x,3) for k,x in d['roofline']['kernel_ms_per_step'].items()}, d['parity']['counters_equal'], d['parity']['records_equal'], 'alt', d.get('alt_path',{}).get('value'))
except Exception as e:
    print('$name cfg $c FAILED', e)
PY
}
for c in 3 5; do
  run refill64 $c X=1
  run refill80 $c MATCHY_B200_IPTRIE_MINB=3
  run r2f $c MATCHY_B200_VARIANT=1
done
