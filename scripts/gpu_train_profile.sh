#!/bin/bash
# launch list of REINFORCE training steps (PNLow, B = 128) under ncu: kernel names, counts and summed durations --
# evidence that no cuDNN / cuBLAS / batch_norm kernel runs in the step.  Writes gpurun_out/launches_train_step.csv
mkdir -p gpurun_out
timeout 800 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train_raw.csv \
  python scripts/bench_train.py --impl own --steps 1 > gpurun_out/ncu_train.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/launches_train_raw.csv")) if len(r) > 10]
hdr = rows[0]; ik = hdr.index("Kernel Name"); iv = hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    k = r[ik]
    c, t = agg.get(k, (0, 0.0))
    agg[k] = (c + 1, t + float(r[iv].replace(",", "")))
with open("gpurun_out/launches_train_step.csv", "w") as f:
    f.write("kernel,launches (4 steps: 3 warm-up + 1 timed),total_duration_ns\n")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"\"{k}\",{c},{t:.0f}\n")
bad = [k for k in agg if any(s in k.lower() for s in ("cudnn", "cublas", "gemm_", "batch_norm", "sgemm", "cutlass")) and "gnnpn" not in k]
print("kernels:", len(agg), "library (cudnn/cublas/batch_norm) kernels:", bad)
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
    print(c, round(t / 1e3), "us", k[:110])
PY
