"""Workload for the ncu captures of the ML-stage kernels (aggregation, node transform, CSR build):

    ncu --set full --clock-control none --import-source on -k regex:'spmm_csr_kernel|tc_mainloop_kernel|gemm_ffma' \
        -o gpurun_out/ml_full -f python scripts/ncu_ml_kernels.py

Aggregation: E = 2^28 edges, mean degree 16, F = 64 and 256, weighted (the GCN form, modelML.py:153) -- far beyond L2.
Node transform: the GCN layer's X.W at scale-up size (B.S = 2^20 service rows, 256 -> 256, modelML.py:98-104) with
bias + eval-BatchNorm + ReLU folded into the epilogue.
Prints the CUDA-event time of each op (outside ncu these are the bench numbers; under ncu they are not)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnnpn_sc_b200 import ops
from scripts.bench_agg import make_csr

dev = torch.device("cuda")


def timed(fn, iters=5):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


out = []
for F in (64, 256):
    N = (1 << 28) // 16
    rowptr, col, val, E = make_csr(N, 16, 0.0, True, dev)
    x = torch.empty(N, F, device=dev).uniform_(-1, 1)
    y = torch.empty(N, F, device=dev)
    ms = timed(lambda: ops.spmm_csr(rowptr, col, val, x, out=y))
    alg = E * (8 + 4 * F) + N * 4 * F + (N + 1) * 8
    out.append({"kernel": "spmm_csr", "E": E, "N": N, "F": F, "ms": ms, "algorithmic_bytes": alg, "alg_GBps": alg / ms / 1e6})
    del x, y, rowptr, col, val
    torch.cuda.empty_cache()

M, Kd, Nd = 1 << 20, 256, 256
a = torch.empty(M, Kd, device=dev).uniform_(-1, 1)
w = torch.empty(Nd, Kd, device=dev).uniform_(-0.1, 0.1)
bias = torch.zeros(Nd, device=dev); scale = torch.ones(Nd, device=dev); shift = torch.zeros(Nd, device=dev)
o = torch.empty(M, Nd, device=dev)
ms = timed(lambda: ops.gemm_bias_act(a, w, bias, scale, shift, "relu", out=o))
out.append({"kernel": "gemm_bias_act (tcgen05 3xTF32)", "M": M, "K": Kd, "N": Nd, "ms": ms,
            "algorithmic_tflops": 2 * M * Kd * Nd / ms / 1e9})
for r in out:
    print(json.dumps(r))
