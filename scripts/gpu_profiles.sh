#!/bin/bash
# end-of-round evidence: ncu launch list of the bench command, full captures of the scan kernels (pair + column-split)
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lstm_seq_kernel -c 2 -o gpurun_out/seq_full -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstm_colsplit_kernel --launch-skip 5 -c 2 -o gpurun_out/colsplit_full -f python scripts/diag_colsplit.py --time-only > gpurun_out/ncu_colsplit.log 2>&1
ls -la gpurun_out | tail -8
