# quick check: parity tests + cfg2 bench (no e2e / cpu baseline)
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
for c in 2 1 3 4; do timeout 300 python bench.py --config $c --gb ${GB:-8} --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_c$c.json 2> gpurun_out/bench_c$c.err; python -c "
import json
d=json.load(open('gpurun_out/bench_c$c.json')); print($c, round(d['value'],1), round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['roofline']['kernel_ms_per_step'].items()}, d['counters'])"; tail -2 gpurun_out/bench_c$c.err; done
