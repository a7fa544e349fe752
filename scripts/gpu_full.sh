#!/bin/bash
# full GPU pass: all parity tests, smoke, default bench (+ reference arm), wait-cycle counters
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -1 gpurun_out/bench_default.json | cut -c1-600
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>&1; tail -1 gpurun_out/bench_reference.json | cut -c1-300
GNNPN_SEQ_PROF=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_prof.log 2>&1; grep "seq prof" gpurun_out/bench_prof.log | tail -2
