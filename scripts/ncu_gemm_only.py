"""ncu workload: the node-transform GEMM alone (see scripts/ncu_ml_kernels.py for the full ML-stage set)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnnpn_sc_b200 import ops
dev = torch.device("cuda")
M = int(os.environ.get("GEMM_M", 1 << 20)); Kd = int(os.environ.get("GEMM_K", 256)); Nd = int(os.environ.get("GEMM_N", 256))
a = torch.empty(M, Kd, device=dev).uniform_(-1, 1)
w = torch.empty(Nd, Kd, device=dev).uniform_(-0.1, 0.1)
bias = torch.zeros(Nd, device=dev); scale = torch.ones(Nd, device=dev); shift = torch.zeros(Nd, device=dev)
o = torch.empty(M, Nd, device=dev)
ts = []
for i in range(6):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.gemm_bias_act(a, w, bias, scale, shift, "relu", out=o); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ms = sorted(ts[1:])[len(ts[1:]) // 2]
ref = torch.relu(a[:4096].double() @ w.double().T).float()
err = float((o[:4096] - ref).abs().max() / ref.abs().max())
print(json.dumps({"kernel": "gemm_bias_act", "M": M, "K": Kd, "N": Nd, "ms": ms, "algorithmic_tflops": 2 * M * Kd * Nd / ms / 1e9,
                  "algorithmic_bytes": 4 * M * (Kd + Nd), "alg_GBps": 4 * M * (Kd + Nd) / ms / 1e6, "max_rel_err_vs_fp64": err}))
