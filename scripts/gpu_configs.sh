#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/configs.jsonl
timeout 900 python scripts/bench_configs.py ml --out gpurun_out/configs.jsonl 2>&1 | tail -8
timeout 900 python scripts/bench_configs.py pipeline --out gpurun_out/configs.jsonl 2>&1 | tail -3
timeout 900 python scripts/bench_configs.py pipeline --instances 3840 --out gpurun_out/configs.jsonl 2>&1 | tail -1
timeout 900 python scripts/bench_configs.py pipeline --instances 1920 --out gpurun_out/configs.jsonl 2>&1 | tail -1
timeout 900 python scripts/bench_configs.py scaleup --instances 128 --iters 2 --out gpurun_out/configs.jsonl 2>&1 | tail -3
timeout 900 python scripts/bench_configs.py scaleup --instances 512 --iters 2 --out gpurun_out/configs.jsonl 2>&1 | tail -3
GNNPN_COLSPLIT=0 timeout 900 python scripts/bench_configs.py pipeline --instances 1920 2>&1 | tail -1
