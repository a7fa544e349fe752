#!/bin/bash
# data-parallel REINFORCE steps / s on 1 / 2 / 4 / 8 GPUs (B = 128 per rank, one flat-bucket NCCL all-reduce per step)
mkdir -p gpurun_out
: > gpurun_out/train_dp.jsonl
timeout 200 python scripts/bench_train.py --impl own --steps 20 --out gpurun_out/train_dp.jsonl > /dev/null 2> gpurun_out/train_g1.err
timeout 200 python scripts/bench_train.py --impl torch --steps 20 --out gpurun_out/train_dp.jsonl > /dev/null 2>> gpurun_out/train_g1.err
timeout 200 python scripts/bench_train.py --impl own --level high --steps 20 --out gpurun_out/train_dp.jsonl > /dev/null 2>> gpurun_out/train_g1.err
for N in 2 4 8; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + N)) \
    scripts/bench_train.py --impl own --steps 20 --out gpurun_out/train_dp.jsonl > /dev/null 2> gpurun_out/train_g$N.err
done
cat gpurun_out/train_dp.jsonl
