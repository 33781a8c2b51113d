TAG=${1:-r1x}
B="python bench.py --gb 0.5 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --chunk-mb 512"
$B > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'tokenize|token_kernel|exact|iptrie' -c 4 -f -o gpurun_out/prof_$TAG $B > gpurun_out/ncu.log 2>&1; tail -2 gpurun_out/ncu.log
