# one GPU round: parity tests, cfg2 bench at several chunk sizes, short runs of the other configs, ncu capture
TAG=${1:-r1x}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
for cm in 512 1024 2040; do
timeout 300 python bench.py --gb 8 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --chunk-mb $cm > gpurun_out/bench_c2_$cm.json 2> gpurun_out/bench_c2_$cm.err; python -c "
import json
d=json.load(open('gpurun_out/bench_c2_$cm.json')); print('cfg2 chunk $cm', round(d['value'],1), round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['roofline']['kernel_ms_per_step'].items()}, d['gpu_launches'])"; tail -2 gpurun_out/bench_c2_$cm.err; done
for c in 1 3 4; do timeout 300 python bench.py --config $c --gb 2 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_c$c.json 2> gpurun_out/bench_c$c.err; python -c "
import json
d=json.load(open('gpurun_out/bench_c$c.json')); print($c, round(d['value'],1), {k:round(v,3) for k,v in d['roofline']['kernel_ms_per_step'].items()}, d['counters'])"; tail -2 gpurun_out/bench_c$c.err; done
B="python bench.py --gb 0.5 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
$B > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'tokenize|token_kernel|exact|iptrie' -c 8 -f -o gpurun_out/prof_$TAG $B > gpurun_out/ncu.log 2>&1; tail -2 gpurun_out/ncu.log
