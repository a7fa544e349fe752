#!/bin/bash
# round-2 refresh: GPU tests (twice: the bench-parity flake watch), default bench line + reference arm, batch sweep,
# REINFORCE step launch list + section times, ncu --set full of the training kernels
mkdir -p gpurun_out
for i in 1 2; do
  timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu_$i.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$i.log
  grep -E "^  logits_|passed|failed|pytest rc" gpurun_out/pytest_gpu_$i.log | cut -c1-300 | tail -6
done
cp gpurun_out/pytest_gpu_2.log gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_r02e.json 2> gpurun_out/bench_r02e.err
tail -1 gpurun_out/bench_r02e.json | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'])
print('enc',d['roofline']['avg_launch_ms'],d['roofline']['frac'],'dec',d['roofline_decode']['avg_launch_ms'],d['roofline_decode']['frac'])
print('clocks',d['clocks'])
for p in d['pipeline'] or []: print(p['workload'],p['value'],p['e2e']['value'],p['stage_ms'])
"
timeout 300 python bench.py --impl reference > gpurun_out/bench_r02e_reference.json 2>> gpurun_out/bench_r02e.err; cut -c1-200 gpurun_out/bench_r02e_reference.json
timeout 600 python scripts/bench_sweep.py --out gpurun_out/pn_batch_sweep_r02e.jsonl > gpurun_out/sweep.log 2>&1
cut -c1-90 gpurun_out/pn_batch_sweep_r02e.jsonl
timeout 300 python scripts/prof_train_step.py 2>&1 | tail -1
bash scripts/gpu_train_profile.sh 2>&1 | tail -18
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"lstm_bptt_cluster_kernel|lstm_colsplit_kernel" -c 4 \
  -o gpurun_out/r02_train_kernels_full -f python scripts/bench_train.py --impl own --steps 1 > gpurun_out/ncu_train_full.log 2>&1
tail -2 gpurun_out/ncu_train_full.log
