#!/bin/bash
# ncu evidence for the ML-stage kernels: timings outside the profiler, launch list, then one full capture
mkdir -p gpurun_out
timeout 600 python scripts/ncu_ml_kernels.py > gpurun_out/ml_kernels_timing.jsonl 2> gpurun_out/ml_kernels_timing.err; cat gpurun_out/ml_kernels_timing.jsonl
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'spmm_csr_kernel|tc_mainloop_kernel|split_tf32' --launch-skip 0 -c 12 -o gpurun_out/ml_full -f python scripts/ncu_ml_kernels.py > gpurun_out/ncu_ml.log 2>&1
tail -3 gpurun_out/ncu_ml.log; ls -la gpurun_out
