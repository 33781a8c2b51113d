for cm in 768 1024 1536 2040; do
timeout 300 python bench.py --gb 10 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --chunk-mb $cm > gpurun_out/bench_sw_$cm.json 2> gpurun_out/bench_sw_$cm.err; python -c "
import json
d=json.load(open('gpurun_out/bench_sw_$cm.json')); print('chunk $cm', round(d['value'],1), round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['roofline']['kernel_ms_per_step'].items()}, d['gpu_launches'])"; tail -2 gpurun_out/bench_sw_$cm.err; done
