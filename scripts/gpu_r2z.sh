#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "fullsize or parity" 2>&1 | tail -4
timeout 400 python bench.py > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2z_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], 'wall', d['value_wall'], d['wall_s_timed_region']/d['steps']*1000, d['e2e']['value'], d['roofline']['kernel_ms_per_step'], d['parity']['counters_equal'], d['parity']['records_equal'])
for k,v in d['per_config'].items(): print(k, v['value'], v['parity']['records_equal'])
PY
bash scripts/gpu_r2_prof.sh r2z
