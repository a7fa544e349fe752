#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lstm_seq_kernel --launch-skip 1 --launch-count 1 -o gpurun_out/dec_full -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_dec.log 2>&1
tail -3 gpurun_out/ncu_dec.log
