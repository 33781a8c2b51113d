#!/bin/bash
# Round-2 closing run: whole GPU suite, the default bench line (as the driver runs it), the launch list of whole steps and one
# --set full capture of iptrie_kernel on config 3 now that its records stay in HBM.  Every ncu command runs after the same
# command has exited 0 without ncu.
mkdir -p gpurun_out
timeout 700 python -m pytest tests -x -q -m gpu > gpurun_out/r2z_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/r2z_tests.log
timeout 600 python bench.py > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; echo "bench exit $?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r2z_bench.json').read().strip().splitlines()[-1])
print('cfg2', round(d['value'],1), 'wall', round(d['value_wall'],1), 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'], d['parity'])
for k,v in (d.get('per_config') or {}).items(): print(k, round(v['value'],1), 'wall', round(v['value_wall'],1), v['dominant_kernel'], round(v['dominant_kernel_frac'],3), v['parity']['counters_equal'], v['parity']['records_equal'])
print('alt', d.get('alt_path'))
print('launches', d.get('gpu_launches'), d.get('clocks'))
PY
BL="python bench.py --gb 4 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-parity --no-per-config"
$BL > gpurun_out/plain_r2z_l.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2z.csv $BL > gpurun_out/ncu_r2z_l.log 2>&1; tail -1 gpurun_out/ncu_r2z_l.log
B3="python bench.py --config 3 --gb 0.5 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-parity --no-per-config --chunk-mb 512"
$B3 > gpurun_out/plain_r2z_c3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'iptrie_kernel|sort_keys|sort_gather|DeviceRadixSort' -c 12 -f -o gpurun_out/prof_r2z_c3 $B3 > gpurun_out/ncu_r2z_c3.log 2>&1; tail -1 gpurun_out/ncu_r2z_c3.log
ncu -i gpurun_out/prof_r2z_c3.ncu-rep --page raw --csv > gpurun_out/prof_r2z_c3_raw.csv 2>/dev/null
rm -f gpurun_out/prof_r2z_c3.ncu-rep
