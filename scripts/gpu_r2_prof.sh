#!/bin/bash
# Round-2 ncu evidence of the code as committed: (1) --set full of the default kernel pair + lookups on one 500 MB piece of
# config 2, (2) the same for config 3 (IP-trie heavy), (3) scan_kernel (MATCHY_B200_FUSED=1) on config 2, (4) launch list of
# whole steps.  Every ncu command runs only after the same command exited 0 without ncu.
mkdir -p gpurun_out
TAG=${1:-r2z}
B="python bench.py --gb 0.5 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-parity --no-per-config --chunk-mb 512"
$B > gpurun_out/plain_$TAG.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'tokenize_kernel|token_kernel|exact_kernel|iptrie_kernel' -c 6 -f -o gpurun_out/prof_${TAG}_c2 $B > gpurun_out/ncu_${TAG}_c2.log 2>&1; tail -2 gpurun_out/ncu_${TAG}_c2.log
B3="python bench.py --config 3 --gb 0.5 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-parity --no-per-config --chunk-mb 512"
$B3 > gpurun_out/plain_${TAG}_c3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'tokenize_kernel|token_kernel|iptrie_kernel' -c 6 -f -o gpurun_out/prof_${TAG}_c3 $B3 > gpurun_out/ncu_${TAG}_c3.log 2>&1; tail -2 gpurun_out/ncu_${TAG}_c3.log
MATCHY_B200_FUSED=1 $B > gpurun_out/plain_${TAG}_f.log 2>&1 && MATCHY_B200_FUSED=1 ncu --set full --clock-control none --import-source on -k regex:'scan_kernel' -c 2 -f -o gpurun_out/prof_${TAG}_fused $B > gpurun_out/ncu_${TAG}_f.log 2>&1; tail -2 gpurun_out/ncu_${TAG}_f.log
BL="python bench.py --gb 4 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-parity --no-per-config"
$BL > gpurun_out/plain_${TAG}_l.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $BL > gpurun_out/ncu_${TAG}_l.log 2>&1; tail -2 gpurun_out/ncu_${TAG}_l.log
# the reports themselves (25 MB each) would push gpurun_out/ past what travels back: keep the raw and source pages as CSV
for r in c2 c3 fused; do
  ncu -i gpurun_out/prof_${TAG}_$r.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_${r}_raw.csv 2>/dev/null
  ncu -i gpurun_out/prof_${TAG}_$r.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/prof_${TAG}_${r}_source.csv 2>/dev/null
  python profiles/ncu_lines.py gpurun_out/prof_${TAG}_${r}_source.csv 60 > gpurun_out/prof_${TAG}_${r}_lines.txt 2>/dev/null
  gzip -f gpurun_out/prof_${TAG}_${r}_source.csv
  rm -f gpurun_out/prof_${TAG}_$r.ncu-rep
done
ls -la gpurun_out | grep $TAG
