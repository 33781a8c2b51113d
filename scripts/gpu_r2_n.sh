#!/bin/bash
# multi-GPU bench exactly as the driver launches it: N ranks over NCCL, config 5, 200 GB total, strong scaling
N=${1:-2}; TAG=${2:-r2n}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/${TAG}_n$N.json 2> gpurun_out/${TAG}_n$N.err; echo "exit $?"
tail -3 gpurun_out/${TAG}_n$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_n$N.json').read().strip().splitlines()[-1])
print('N=$N', d['config']['workload'][:40], 'value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],2), 'wall', round(d['value_wall'],1), 'e2e', d['e2e'] and round(d['e2e']['value'],1), 'parity', d['parity'] and (d['parity']['counters_equal'], d['parity']['records_equal']), d['roofline']['kernel_ms_per_step'], d.get('host_us_last_step'), d['counters']['matches'], d['config'].get('generated'))
PY
