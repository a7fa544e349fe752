"""One in-degree-Zipf aggregation call per feature width (for an ncu launch list of the split path's kernels)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnnpn_sc_b200 import ops
from scripts.bench_agg import make_csr_indegree_zipf, make_csr
dev = torch.device("cuda")
E_t = 1 << 28
N = E_t // 16
for F in (32, 128):
    rowptr, col, val, E, dmax = make_csr_indegree_zipf(N, E_t, True, dev, floor=8)
    x = torch.empty(N, F, device=dev).uniform_(-1, 1)
    y = torch.empty(N, F, device=dev)
    for _ in range(2):
        ops.spmm_csr(rowptr, col, val, x, out=y)
    torch.cuda.synchronize()
    del rowptr, col, val
    rowptr, col, val, E = make_csr(N, 8, 0.0, True, dev)
    for _ in range(2):
        ops.spmm_csr(rowptr, col, val, x, out=y, long_row_threshold=0)
    torch.cuda.synchronize()
    del rowptr, col, val, x, y
    torch.cuda.empty_cache()
