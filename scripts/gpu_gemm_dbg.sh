#!/bin/bash
mkdir -p gpurun_out
for shape in "65536 24 256" "1048576 128 128"; do set -- $shape
    echo "M=$1 K=$2 N=$3 tma-store"; GEMM_M=$1 GEMM_K=$2 GEMM_N=$3 timeout 25 python scripts/ncu_gemm_only.py 2>&1 | tail -1 | cut -c1-200
done 2>&1 | tee gpurun_out/gemm_dbg.log
echo synccheck; GEMM_M=16384 GEMM_K=24 GEMM_N=256 timeout 120 compute-sanitizer --tool synccheck python scripts/ncu_gemm_only.py 2>&1 | tail -15 | cut -c1-220 | tee -a gpurun_out/gemm_dbg.log
echo memcheck; GEMM_M=16384 GEMM_K=24 GEMM_N=256 timeout 120 compute-sanitizer --tool memcheck python scripts/ncu_gemm_only.py 2>&1 | tail -15 | cut -c1-220 | tee -a gpurun_out/gemm_dbg.log
