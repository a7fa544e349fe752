#!/bin/bash
# node-transform GEMM: parity tests, timing at the scale-up shape, ncu full capture
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_graph_gpu.py tests/test_ml_gpu.py -m gpu -x -q -s 2>&1 | grep -v "^$" | tail -40 > gpurun_out/pytest_gemm.log; tail -30 gpurun_out/pytest_gemm.log
for shape in "1048576 256 256" "1048576 24 256" "1048576 128 128" "200000 256 128"; do set -- $shape; GEMM_M=$1 GEMM_K=$2 GEMM_N=$3 timeout 120 python scripts/ncu_gemm_only.py; done 2>&1 | tee gpurun_out/gemm_timing.jsonl
GNNPN_GEMM_V1=1 timeout 120 python scripts/ncu_gemm_only.py 2>&1 | tail -1 | tee gpurun_out/gemm_timing_v1.jsonl
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"node_transform_kernel" -c 2 -o gpurun_out/gemm_v2 -f python scripts/ncu_gemm_only.py > gpurun_out/ncu_gemm2.log 2>&1; tail -2 gpurun_out/ncu_gemm2.log
