#!/bin/bash
# N-GPU weak-scaling point of the bench under torchrun (instance-sharded, no collective): bash scripts/gpu_multi_n.sh N
N=${1:-4}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29527 bench.py --gpus $N --steps 5 --warmup 3 --no-aggregation > gpurun_out/bench_g$N.json 2> gpurun_out/bench_g$N.err
tail -1 gpurun_out/bench_g$N.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('gpus',d['n_gpus'],'value',round(d['value']),'e2e',round(d['e2e']['value']),'ms',d['ms_per_step'], d['clocks'])"
