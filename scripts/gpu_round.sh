timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --gb 4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; cat gpurun_out/bench_c2.json; tail -3 gpurun_out/bench_c2.err
for c in 1 3 4; do timeout 300 python bench.py --config $c --gb 2 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_c$c.json 2> gpurun_out/bench_c$c.err; python -c "
import json,sys
d=json.load(open('gpurun_out/bench_c$c.json')); print($c, d['value'], d['roofline']['kernel_ms_per_step'], d['counters'])"; tail -2 gpurun_out/bench_c$c.err; done
B="python bench.py --gb 0.5 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
$B > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'tokenize|token_kernel|exact|iptrie' -c 8 -f -o gpurun_out/prof_r1m $B > gpurun_out/ncu.log 2>&1; tail -2 gpurun_out/ncu.log
