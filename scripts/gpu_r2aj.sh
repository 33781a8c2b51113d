#!/bin/bash
# (historical: the parked / refill iptrie_kernel variants these runs compared were measured and then removed — profiles/README.md "last session"; MATCHY_B200_VARIANT=1 and MATCHY_B200_IPTRIE_MINB=3 select nothing in the committed library)
# r2aj: stand-alone kernel times (MATCHY_B200_SERIAL=1: lookups on the compute stream, nothing overlaps) of the refill iptrie kernel vs r2f on config 3,
# then ncu --set full of the refill kernel on one 500 MB piece
mkdir -p gpurun_out
run() {
  local name=$1 c=$2; shift 2
  env "$@" timeout 300 python bench.py --config $c --gb 8 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-per-config --no-parity > gpurun_out/r2aj_c${c}_$name.json 2> gpurun_out/r2aj_c${c}_$name.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2aj_c${c}_$name.json').read().strip().splitlines()[-1])
    print('$name cfg $c', round(d['value'],1), round(d['ms_per_step'],3), {k:round(x,3) for k,x in d['roofline']['kernel_ms_per_step'].items()})
except Exception as e:
    print('$name cfg $c FAILED', e)
PY
}
run serial_refill80 3 MATCHY_B200_SERIAL=1 MATCHY_B200_IPTRIE_MINB=3
run serial_refill64 3 MATCHY_B200_SERIAL=1
run serial_r2f 3 MATCHY_B200_SERIAL=1 MATCHY_B200_VARIANT=1
B3="python bench.py --config 3 --gb 0.5 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-parity --no-per-config --chunk-mb 512"
MATCHY_B200_IPTRIE_MINB=3 ncu --set full --clock-control none --import-source on -k regex:'iptrie_kernel' -c 2 -f -o gpurun_out/prof_r2aj_c3 $B3 > gpurun_out/ncu_r2aj_c3.log 2>&1; tail -2 gpurun_out/ncu_r2aj_c3.log
ncu -i gpurun_out/prof_r2aj_c3.ncu-rep --page raw --csv > gpurun_out/prof_r2aj_c3_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_r2aj_c3.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/prof_r2aj_c3_source.csv 2>/dev/null
python profiles/ncu_lines.py gpurun_out/prof_r2aj_c3_source.csv 60 > gpurun_out/prof_r2aj_c3_lines.txt 2>/dev/null
gzip -f gpurun_out/prof_r2aj_c3_source.csv
rm -f gpurun_out/prof_r2aj_c3.ncu-rep
