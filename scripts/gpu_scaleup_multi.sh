#!/bin/bash
# BASELINE config 4 (100 tasks x 1000 candidates) instance-sharded over 1 / 2 / 4 / 8 GPUs of one node, no collective.
# usage (under gpurun --gpus 8): bash scripts/gpu_scaleup_multi.sh
mkdir -p gpurun_out
: > gpurun_out/scaleup_multi.jsonl
timeout 300 python bench.py --workload scaleup --steps 2 --warmup 1 2> gpurun_out/scaleup_g1.err | tail -1 >> gpurun_out/scaleup_multi.jsonl
for N in 2 4 8; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) \
    bench.py --gpus $N --workload scaleup --steps 2 --warmup 1 2> gpurun_out/scaleup_g$N.err | tail -1 >> gpurun_out/scaleup_multi.jsonl
done
python - <<'PY'
import json
rows = [json.loads(l) for l in open("gpurun_out/scaleup_multi.jsonl") if l.strip().startswith("{")]
base = rows[0]["value"] if rows else None
for r in rows:
    print(r["n_gpus"], "GPUs:", round(r["value"], 1), "inst/s, per GPU", round(r["per_gpu_instances_per_s"], 1),
          "efficiency", round(r["value"] / (base * r["n_gpus"]), 4), "e2e", round(r["e2e"]["value"], 1), "ms/step", round(r["ms_per_step"], 1))
PY
