#!/bin/bash
# profiling experiments for the tcgen05 LSTM step (GNNPN_TC_DBG knobs), one JSON line each
for d in 0 1 2; do
  echo "== GNNPN_TC_DBG=$d"
  GNNPN_TC_DBG=$d timeout 300 python bench.py --steps 3 --warmup 2 --instances 18944 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['avg_launch_ms'])"
done
