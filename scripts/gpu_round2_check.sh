#!/bin/bash
# round-2 check pass: GPU tests, the default bench line, the batch sweep through pipeline.low_high
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_r02d.json 2> gpurun_out/bench_r02d.err
tail -1 gpurun_out/bench_r02d.json | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'])
print('enc',d['roofline']['avg_launch_ms'],d['roofline']['frac'],'dec',d['roofline_decode']['avg_launch_ms'],d['roofline_decode']['frac'])
print('clocks',d['clocks'])
for p in d['pipeline'] or []: print(p['workload'],p['value'],p['e2e']['value'],p['stage_ms'])
"
tail -3 gpurun_out/bench_r02d.err
timeout 600 python scripts/bench_sweep.py --out gpurun_out/pn_batch_sweep_r02d.jsonl > gpurun_out/sweep.log 2>&1
cut -c1-100 gpurun_out/pn_batch_sweep_r02d.jsonl
