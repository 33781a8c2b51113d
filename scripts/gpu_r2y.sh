#!/bin/bash
# r2y: IP-less databases skip the IP token list; suite + bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2y_tests.log; cat gpurun_out/r2y_tests.log
timeout 400 python bench.py > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2y_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['value_wall'], d['e2e']['value'], d['roofline']['kernel_ms_per_step'], d['parity']['counters_equal'], d['parity']['records_equal'])
for k,v in d['per_config'].items(): print(k, v['value'], v['kernel_ms_per_step'], v['parity'])
print(d['alt_path']['value'])
PY
