#!/bin/bash
for cfg in "2 0" "3 2" "3 4" "3 6"; do
  set -- $cfg
  echo "== GNNPN_SEQ_DEC=$1 PF=$2"
  GNNPN_SEQ_DEC=$1 GNNPN_SEQ_PF=$2 GNNPN_SEQ_PROF=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline 2>&1 | grep "seq prof dec" | tail -1
  GNNPN_SEQ_DEC=$1 GNNPN_SEQ_PF=$2 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value',d['value'],'ms',d['ms_per_step'])"
done
