#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_pn_gpu.py -m gpu -x -q 2>&1 | tail -3
for f in 0 1 2 3; do
  echo "== GNNPN_SEQ_DEC=$f"
  GNNPN_SEQ_DEC=$f GNNPN_SEQ_PROF=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline 2>&1 | grep "seq prof dec" | tail -1
  GNNPN_SEQ_DEC=$f timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value',d['value'],'ms',d['ms_per_step'])"
done
