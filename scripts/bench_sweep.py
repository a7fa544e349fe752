"""Batch-size sweep of the greedy PNLow -> PNHigh decode (n from 128 to 2^17), QWS shape, inputs
resident in HBM, CUDA events, median of 5 after 2 warm-ups.  Goes through the module path the pipeline uses
(gnnpn_sc_b200.pipeline.low_high: the two encoders on two streams, then the decoders), so mid-size batches that leave SMs
free run both encoders concurrently.  One JSON line per n: instances/s, ms, and which scan the dispatcher used.

    python scripts/bench_sweep.py [--out gpurun_out/pn_batch_sweep.jsonl] [--sequential]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnnpn_sc_b200 import modelPN as M, ops
from gnnpn_sc_b200.pipeline import low_high
from gnnpn_sc_b200.synth import pn_instances
from gnnpn_sc_b200.weights import reference_shaped_state_dict

K, N, H, F = 47, 5, 256, 8
L = K * N
ap = argparse.ArgumentParser()
ap.add_argument("--out", default="gpurun_out/pn_batch_sweep.jsonl")
ap.add_argument("--sizes", default="128,512,1024,1920,2048,3072,3840,4096,6144,8192,9472,12288,18944,37888,75776")
ap.add_argument("--sequential", action="store_true", help="one stream (encoders back to back) for comparison")
a = ap.parse_args()
dev = torch.device("cuda")
nets = []
for level, seed in (("Low", 1), ("High", 2)):
    m = M.CombinatorialRL(0, H, L, 0, 10, 1, M.reward, "Dot", N, K, level=level)
    m.load_state_dict(reference_shaped_state_dict(H, F, seed))
    m.actor.check_inputs = False                                   # no host sync inside the timed loop
    nets.append(m.to(dev).eval())
low, high = nets
side = torch.cuda.Stream(dev)
os.makedirs(os.path.dirname(a.out), exist_ok=True)
with open(a.out, "w") as f:
    for n in [int(s) for s in a.sizes.split(",")]:
        x = pn_instances(n, K, N, seed=3).to(dev)

        def step():
            with torch.no_grad():
                if a.sequential:
                    _, _, _, _, lat = low(x, None, sample="greedy", training="SL")
                    R, _, _, idx, _ = high(x, None, lat, sample="greedy", training="RL")
                else:
                    _, R, idx = low_high(low, high, x, side)
            return R

        for _ in range(2):
            step()
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); step(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[2]
        lay = low.actor.last["enc_layout"]
        line = {"n": n, "ms": ms, "instances_per_s": n / ms * 1e3, "groups_of_128": (n + 127) // 128,
                "scan": "cta-pair (blocked encodings, fused pointer dots)" if lay == ops.ENC_BLOCKED128 else "column-split",
                "encoders": "sequential" if a.sequential else "two streams"}
        print(json.dumps(line), flush=True)
        f.write(json.dumps(line) + "\n")
        del x
        low.actor.last = high.actor.last = None
        torch.cuda.empty_cache()
