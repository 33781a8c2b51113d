mkdir -p gpurun_out
TAG=${1:-r2i}
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest.log 2>&1; tail -6 gpurun_out/${TAG}_pytest.log
B="python bench.py --gb 0.5 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-parity --chunk-mb 512"
$B > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'scan_kernel|token_kernel|exact|iptrie' -c 4 -f -o gpurun_out/prof_$TAG $B > gpurun_out/ncu.log 2>&1; tail -2 gpurun_out/ncu.log
ncu -i gpurun_out/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_$TAG.ncu-rep --page source --csv -k regex:scan_kernel > gpurun_out/prof_${TAG}_source.csv 2>/dev/null; ls -la gpurun_out/prof_${TAG}*
