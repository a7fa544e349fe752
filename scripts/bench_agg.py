"""Aggregation micro-benchmark (BASELINE.json configs[4]): CSR segment-reduce
GB/s over edge count and feature width against the measured HBM peak.

    python scripts/bench_agg.py [--quick] [--out gpurun_out/agg_sweep.jsonl]

ALGORITHMIC bytes = E*(4 + 4w + 4F) + N*4F + (N+1)*8   (col, val, one gathered row per edge; output
written once; int64 rowptr read once).  Uniform in-degree d (16, or 64 when N*F would not fit);
sources uniform or Zipf-like skewed.  Points whose feature matrix fits the 126 MB L2 are labelled
L2-resident (gathers are served by L2, the figure may exceed HBM peak)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnnpn_sc_b200 import ops

def peak():
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    return json.load(open(p))["hbm_gbs"] if os.path.exists(p) else 6650.0

def make_csr(N, d, skew, weighted, dev):
    E = N * d
    g = torch.Generator(device=dev).manual_seed(7)
    rowptr = torch.arange(0, E + 1, d, device=dev, dtype=torch.int64)
    if skew > 0:
        col = torch.empty(E, device=dev, dtype=torch.int32)
        chunk = 1 << 26
        for s in range(0, E, chunk):
            m = min(chunk, E - s)
            u = torch.rand(m, device=dev, generator=g)
            col[s:s + m] = (N * u.pow(1.0 + skew)).to(torch.int32).clamp_(max=N - 1)
    else:
        col = torch.randint(0, N, (E,), device=dev, dtype=torch.int32, generator=g)
    val = None
    if weighted:
        val = torch.rand(E, device=dev, generator=g)
    return rowptr, col, val, E

def run_point(E_target, F, skew, weighted, iters, flush):
    dev = torch.device("cuda")
    d = 16
    N = E_target // d
    if 2 * N * F * 4 > 120e9:
        d = 64
        N = E_target // d
    rowptr, col, val, E = make_csr(N, d, skew, weighted, dev)
    x = torch.empty(N, F, device=dev).uniform_(-1, 1)
    y = torch.empty(N, F, device=dev)
    for _ in range(3):
        ops.spmm_csr(rowptr, col, val, x, out=y)
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.add_(1.0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ops.spmm_csr(rowptr, col, val, x, out=y); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = sorted(ts)[len(ts) // 2]
    w = 1 if weighted else 0
    alg = E * (4 + 4 * w + 4 * F) + N * 4 * F + (N + 1) * 8
    compulsory = E * (4 + 4 * w) + 2 * N * 4 * F + (N + 1) * 8
    pk = peak()
    return {"E": E, "N": N, "deg": d, "F": F, "weighted": bool(weighted), "skew": skew, "ms": ms,
            "alg_GBps": alg / ms / 1e6, "frac_of_hbm_peak": alg / ms / 1e6 / pk,
            "compulsory_GBps": compulsory / ms / 1e6,
            "l2_resident": bool(N * F * 4 < 100e6), "peak_GBps": pk}

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--out", default="gpurun_out/agg_sweep.jsonl")
    args = ap.parse_args()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8).float() if False else torch.zeros(64 << 20, device="cuda")
    Es = [1 << 20, 1 << 24] if args.quick else [1 << 20, 1 << 22, 1 << 24, 1 << 26, 1 << 28, 1 << 30]
    Fs = [32, 64, 128, 256]
    with open(args.out, "w") as f:
        for E in Es:
            for F in Fs:
                for skew, weighted in ((0.0, True), (0.0, False), (1.0, True)):
                    if (skew > 0 or not weighted) and E not in (1 << 24, 1 << 28):
                        continue
                    try:
                        r = run_point(E, F, skew, weighted, 3 if E >= (1 << 28) else 7, flush)
                    except torch.OutOfMemoryError:
                        r = {"E": E, "F": F, "skew": skew, "weighted": weighted, "error": "oom"}
                    torch.cuda.empty_cache()
                    f.write(json.dumps(r) + "\n"); f.flush()
                    print(json.dumps(r))

if __name__ == "__main__":
    main()
