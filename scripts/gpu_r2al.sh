#!/bin/bash
# r2al: per-piece result sort on the device (cub radix sort of 64-bit keys + gather) before the records' D2H copy, vs the host sort
# (MATCHY_B200_HOST_SORT=1); whole GPU suite first
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/r2al_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/r2al_tests.log
run() {
  local name=$1 c=$2 gb=$3; shift 3
  env "$@" timeout 300 python bench.py --config $c --gb $gb --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-per-config --parity-gb 1 > gpurun_out/r2al_c${c}_$name.json 2> gpurun_out/r2al_c${c}_$name.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2al_c${c}_$name.json').read().strip().splitlines()[-1])
    print('$name cfg $c', round(d['value'],1), round(d['ms_per_step'],3), 'wall', round(d['value_wall'],1), d.get('host_us_last_step'), {k:round(x,3) for k,x in d['roofline']['kernel_ms_per_step'].items()}, d['parity']['counters_equal'], d['parity']['records_equal'])
except Exception as e:
    print('$name cfg $c FAILED', e)
PY
}
run devsort 3 8 X=1
run hostsort 3 8 MATCHY_B200_HOST_SORT=1
run devsort 5 8 X=1
run devsort 2 10 X=1
run devsort 4 8 X=1
