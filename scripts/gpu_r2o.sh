mkdir -p gpurun_out
TAG=${1:-r2o}
export MATCHY_B200_WIDE_SCAN=1 MATCHY_B200_WIDE_WARPS=28 MATCHY_B200_SERIAL=1
for c in 2 3 4 1 5; do
timeout 600 python bench.py --config $c --gb 8 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/${TAG}_ser_c$c.json 2> gpurun_out/${TAG}_ser_c$c.err; python -c "
import json
d=json.load(open('gpurun_out/${TAG}_ser_c$c.json')); print('serial28', $c, round(d['value'],1), round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['roofline']['kernel_ms_per_step'].items()}, d['counters']['matches'])"; tail -2 gpurun_out/${TAG}_ser_c$c.err; done
