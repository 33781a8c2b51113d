timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_e2e.json 2> gpurun_out/bench_e2e.err; python -c "
import json
d=json.load(open('gpurun_out/bench_e2e.json')); print('value', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', d['e2e']['value'], {k:round(v,3) for k,v in d['roofline']['kernel_ms_per_step'].items()})"; tail -2 gpurun_out/bench_e2e.err
