#!/bin/bash
# (historical: the parked / refill iptrie_kernel variants these runs compared were measured and then removed — profiles/README.md "last session"; MATCHY_B200_VARIANT=1 and MATCHY_B200_IPTRIE_MINB=3 select nothing in the committed library)
# r2ak: match records appended in HBM and copied per piece by the copy engine (default) vs stored straight into pinned host memory
# (MATCHY_B200_RECS_ZEROCOPY=1), with the refill and the r2f iptrie kernels; whole GPU suite first
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/r2ak_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/r2ak_tests.log
run() {
  local name=$1 c=$2 gb=$3; shift 3
  env "$@" timeout 300 python bench.py --config $c --gb $gb --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-per-config --parity-gb 1 > gpurun_out/r2ak_c${c}_$name.json 2> gpurun_out/r2ak_c${c}_$name.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2ak_c${c}_$name.json').read().strip().splitlines()[-1])
    print('$name cfg $c', round(d['value'],1), round(d['ms_per_step'],3), 'wall', round(d['value_wall'],1), {k:round(x,3) for k,x in d['roofline']['kernel_ms_per_step'].items()}, d['parity']['counters_equal'], d['parity']['records_equal'])
except Exception as e:
    print('$name cfg $c FAILED', e)
PY
}
run hbm_refill 3 8 X=1
run hbm_r2f 3 8 MATCHY_B200_VARIANT=1
run zc_r2f 3 8 MATCHY_B200_VARIANT=1 MATCHY_B200_RECS_ZEROCOPY=1
run hbm_refill 5 8 X=1
run hbm_r2f 5 8 MATCHY_B200_VARIANT=1
run hbm_refill 2 10 X=1
run hbm_refill 1 4 X=1
