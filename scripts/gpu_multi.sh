#!/bin/bash
# 2-GPU pass: weak-scaling bench under torchrun (instance-sharded, no collective) + DP training check over NCCL
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/smi_multi.txt
for n in 1 2; do
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_g1.json 2> gpurun_out/bench_g1.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/bench_g$n.json 2> gpurun_out/bench_g$n.err
  fi
  tail -1 gpurun_out/bench_g$n.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('gpus',d['n_gpus'],'value',round(d['value']),'e2e',round(d['e2e']['value']),'ms',d['ms_per_step'])"
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 scripts/dp_check.py > gpurun_out/dp_check.log 2>&1; tail -5 gpurun_out/dp_check.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 | tail -1 | cut -c1-200
