mkdir -p gpurun_out
TAG=${1:-r2v}
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest.log 2>&1; tail -5 gpurun_out/${TAG}_pytest.log
( time timeout 1500 python bench.py ) > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -4 gpurun_out/${TAG}_bench.err; python -c "
import json
d=json.load(open('gpurun_out/${TAG}_bench.json'))
print('value', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'wall', round(d['value_wall'],1), 'e2e', d['e2e'], 'cpu', d['cpu_baseline'])
print('roofline', {k:(round(v,4) if isinstance(v,float) else v) for k,v in d['roofline'].items() if k not in ('traffic_source','peak_source')})
print('parity', d['parity'])
for k,v in d['per_config'].items(): print(k, round(v['value'],1), v['dominant_kernel'], round(v['dominant_kernel_frac'],3), v['parity'])
print('alt', d['alt_path']['value'], d['alt_path']['kernel_ms_per_step'])
"
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/${TAG}_ref.json 2> gpurun_out/${TAG}_ref.err; cat gpurun_out/${TAG}_ref.json | cut -c1-600
