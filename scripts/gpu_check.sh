#!/bin/bash
# One GPU-box pass: parity tests, bench (persistent CTA-pair / single-CTA / per-step), ncu launch list + full capture.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -1 gpurun_out/bench_default.json
GNNPN_SEQ_CG=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cg1.json 2>&1; tail -1 gpurun_out/bench_cg1.json
GNNPN_SEQ=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_step.json 2>&1; tail -1 gpurun_out/bench_step.json
GNNPN_SEQ_PROF=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_prof.log 2>&1; grep "seq prof" gpurun_out/bench_prof.log | tail -4
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>&1; tail -1 gpurun_out/bench_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lstm_seq_kernel -c 2 -o gpurun_out/seq_full -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
