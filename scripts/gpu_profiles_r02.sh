#!/bin/bash
# Round-2 ncu evidence (run under gpurun, one GPU).  Outputs under gpurun_out/, summarised into profiles/ by hand.
mkdir -p gpurun_out
# 1. launch list of the bench command (per-launch durations are cold-cache and serialised: shares, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-aggregation --no-pipeline > gpurun_out/ncu_launch_r02.log 2>&1
# 2. full counter set for the blocked encoder scan and the fused decoder at n = 18,944 (third iteration of the timing loop)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstm_seq_kernel --launch-skip 4 -c 2 \
  -o gpurun_out/r02_seq_full -f python scripts/diag_fused.py --time-only > gpurun_out/ncu_seq_r02.log 2>&1
# 3. same for the Normal shape (K = 50, N = 10)
timeout 600 ncu --set full --clock-control none -k regex:lstm_seq_kernel --launch-skip 5 -c 1 \
  -o gpurun_out/r02_dec_normal_full -f python scripts/diag_fused.py --time-only --normal > gpurun_out/ncu_dec_normal_r02.log 2>&1
# 4. ML training step launch list
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_ml_train.csv \
  python scripts/ml_train_launches.py > gpurun_out/ncu_ml_train.log 2>&1
# 5. aggregation sweep (timed) and its DRAM-bytes pass (one launch per point)
timeout 900 python scripts/bench_agg.py --out gpurun_out/agg_sweep_r02.jsonl > gpurun_out/agg_sweep_r02.log 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:spmm --csv \
  --log-file gpurun_out/agg_dram_r02.csv python scripts/bench_agg.py --quick --one-launch --out gpurun_out/agg_one_launch.jsonl > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/r02_* gpurun_out/agg_* | tail -12
tail -3 gpurun_out/ncu_seq_r02.log
