# config 2 only: short bench (kernel times), optionally the parity tests first (TESTS=1)
if [ -n "$TESTS" ]; then timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5; fi
for i in 1 2; do timeout 300 python bench.py --config 2 --gb ${GB:-8} --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; python -c "
import json
d=json.load(open('gpurun_out/bench_c2.json')); print(2, round(d['value'],1), round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['roofline']['kernel_ms_per_step'].items()}, d['clocks'])"; tail -2 gpurun_out/bench_c2.err; done
