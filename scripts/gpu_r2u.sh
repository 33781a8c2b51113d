mkdir -p gpurun_out
TAG=${1:-r2u}
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "device_generator" ) > gpurun_out/${TAG}_pytest.log 2>&1; tail -4 gpurun_out/${TAG}_pytest.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 --total-gb 40 ) > gpurun_out/${TAG}_n2_40.json 2> gpurun_out/${TAG}_n2_40.err; tail -c 1800 gpurun_out/${TAG}_n2_40.json; tail -5 gpurun_out/${TAG}_n2_40.err
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 3 --warmup 3 ) > gpurun_out/${TAG}_n2_200.json 2> gpurun_out/${TAG}_n2_200.err; python -c "
import json
d=json.load(open('gpurun_out/${TAG}_n2_200.json')); print('N=2 200GB', round(d['value'],1), round(d['ms_per_step'],2), d['config']['generated'], d['counters'], d['parity'], d['e2e'])"; tail -5 gpurun_out/${TAG}_n2_200.err
