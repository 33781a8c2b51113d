# ncu --set full of the token / tokenize kernels on one 500 MB piece of config $1 (default 3)
CFG=${1:-3}; TAG=${2:-c$CFG}
B="python bench.py --config $CFG --gb 0.5 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --chunk-mb 512"
$B > gpurun_out/plain_$TAG.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'tokenize|token_kernel|iptrie' -c 3 -f -o gpurun_out/prof_$TAG $B > gpurun_out/ncu_$TAG.log 2>&1; tail -2 gpurun_out/ncu_$TAG.log
