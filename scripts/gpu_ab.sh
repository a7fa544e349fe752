#!/bin/bash
for f in 2 10; do
GNNPN_SEQ_DEC=$f GNNPN_SEQ_PROF=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline 2>&1 | grep -A1 "seq prof dec" | tail -2
done
