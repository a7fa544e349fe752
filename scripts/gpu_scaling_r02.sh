#!/bin/bash
# headline bench on 1 / 2 / 4 / 8 GPUs of one node (weak scaling, no data-path collective), one JSON line per N
mkdir -p gpurun_out
: > gpurun_out/scaling_r02.jsonl
python bench.py --gpus 1 --steps 5 --warmup 3 --no-pipeline --no-aggregation --no-cpu-baseline 2> gpurun_out/scale_g1.err | tail -1 >> gpurun_out/scaling_r02.jsonl
for N in 2 4 8; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700 + N)) \
    bench.py --gpus $N --steps 5 --warmup 3 2> gpurun_out/scale_g$N.err | tail -1 >> gpurun_out/scaling_r02.jsonl
done
python - <<'PY'
import json
rows = [json.loads(l) for l in open("gpurun_out/scaling_r02.jsonl") if l.strip()]
base = rows[0]["value"]
for r in rows:
    print(r["n_gpus"], round(r["value"]), round(r["e2e"]["value"]), round(r["ms_per_step"], 2), "eff", round(r["value"] / (base * r["n_gpus"]), 3),
          "e2e eff", round(r["e2e"]["value"] / (rows[0]["e2e"]["value"] * r["n_gpus"]), 3))
PY
