#!/bin/bash
# ncu evidence for the late-round kernels: two-group column-split encoder, device-resident ESWOA search
mkdir -p gpurun_out
GNNPN_COLSPLIT_G=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstm_colsplit_kernel --launch-skip 24 -c 1 -o gpurun_out/colsplit_g2_full -f python scripts/diag_colsplit.py --time-only > gpurun_out/ncu_colsplit_g2.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:woa_search_kernel -c 1 --launch-skip 1 -o gpurun_out/woa_search_full -f python scripts/bench_woa.py --instances 1024 --cpu-instances 1 > gpurun_out/ncu_woa.log 2>&1
ls -la gpurun_out/*.ncu-rep
