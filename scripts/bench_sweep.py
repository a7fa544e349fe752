"""Batch-size sweep of the greedy PNLow -> PNHigh decode (n from 128 to 2^17), QWS shape, inputs
resident in HBM, CUDA events, median of 5 after 2 warm-ups.  One JSON line per n: instances/s, ms, and which scan the
dispatcher used (column-split cluster scan up to two waves of clusters, CTA-pair scan above).

    python scripts/bench_sweep.py [--out gpurun_out/pn_batch_sweep.jsonl]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnnpn_sc_b200 import modelPN as M, ops
from gnnpn_sc_b200.synth import pn_instances
from gnnpn_sc_b200.weights import reference_shaped_state_dict

K, N, H, F = 47, 5, 256, 8
L = K * N
ap = argparse.ArgumentParser()
ap.add_argument("--out", default="gpurun_out/pn_batch_sweep.jsonl")
ap.add_argument("--sizes", default="128,512,1024,1920,3840,4096,8192,18944,37888,75776,131072")
a = ap.parse_args()
dev = torch.device("cuda")
nets = []
for level, seed in (("Low", 1), ("High", 2)):
    m = M.CombinatorialRL(0, H, L, 0, 10, 1, M.reward, "Dot", N, K, level=level)
    m.load_state_dict(reference_shaped_state_dict(H, F, seed))
    nets.append(m.to(dev).eval())
w = [n.actor._packed_weights() for n in nets]
os.makedirs(os.path.dirname(a.out), exist_ok=True)
with open(a.out, "w") as f:
    for n in [int(s) for s in a.sizes.split(",")]:
        x = pn_instances(n, K, N, seed=3).to(dev)
        c = torch.empty(n, H, device=dev)
        bufs = [(torch.empty(n, K, H, device=dev), torch.empty(K, n, device=dev, dtype=torch.int32),
                 torch.empty(n, L, device=dev), torch.empty(n, L, device=dev)) for _ in range(2)]
        ws = ops.pn_workspace(n, H, dev, "tc")
        lay = ops.pn_enc_layout(n, L, F, K, N, True)
        enc = ops.enc_out_empty(n, L, H, lay, dev)

        def step():
            lat = None
            for lvl, (ew, dw) in enumerate(w):
                ops.lstm_encode(x, ew, H, enc, c, workspace=ws, layout=lay)
                _, idx, lat, _ = ops.pn_decode_greedy(x, enc, c, dw, K, N, latent_win=lat, out=bufs[lvl], workspace=ws,
                                                      enc_layout=lay)
            return ops.pn_reward(x, idx)[2]

        for _ in range(2):
            step()
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); step(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[2]
        groups = (n + 127) // 128
        line = {"n": n, "ms": ms, "instances_per_s": n / ms * 1e3, "scan": "cta-pair (blocked encodings, fused pointer dots)" if lay == ops.ENC_BLOCKED128 else "column-split",
                "groups_of_128": groups}
        print(json.dumps(line), flush=True)
        f.write(json.dumps(line) + "\n")
        del x, enc, c, bufs, ws
        torch.cuda.empty_cache()
