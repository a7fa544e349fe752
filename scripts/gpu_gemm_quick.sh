#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_graph_gpu.py -m gpu -x -q -k "node_transform or gcn_layer" 2>&1 | tail -3
for shape in "1048576 256 256" "1048576 24 256" "1048576 128 128" "200000 256 128"; do set -- $shape; GEMM_M=$1 GEMM_K=$2 GEMM_N=$3 timeout 120 python scripts/ncu_gemm_only.py; done 2>&1 | tee gpurun_out/gemm_timing.jsonl
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"node_transform_kernel" -c 1 -o gpurun_out/gemm_v2 -f python scripts/ncu_gemm_only.py > gpurun_out/ncu_gemm2.log 2>&1; tail -2 gpurun_out/ncu_gemm2.log
GNNPN_GEMM_DIRECT_STORE=1 timeout 120 python scripts/ncu_gemm_only.py 2>&1 | tail -1 | tee gpurun_out/gemm_timing_direct_store.jsonl
