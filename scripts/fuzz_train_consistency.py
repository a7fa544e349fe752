"""Randomised consistency of the REINFORCE gradient paths: fused decode-with-saves on the tensor-core scan + cluster BPTT
(default) against the strict-fp32 decode + FFMA replay + per-step BPTT kernels (impl = "ffma", bptt = 0) on the same
greedy picks, over random (B, K, N) including batches above 1,024.  Exits non-zero when a gradient differs by > 1e-4 of
its tensor's largest entry."""
import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnnpn_sc_b200 import modelPN as M, ops
from gnnpn_sc_b200.synth import pn_instances
from gnnpn_sc_b200.weights import reference_shaped_state_dict

rng = random.Random(int(os.environ.get("FUZZ_SEED", "2")))
bad = 0
for it, B in enumerate([5, 129, 1100, 2100, 40, 777]):
    K = rng.randint(2, 10)
    N = rng.choice([2, 4, 5, 8, 10, 17])
    x = pn_instances(B, K, N, seed=it).cuda()
    high = it % 2 == 1
    lat = [torch.randn(B, K * N, device="cuda") for _ in range(K)] if high else None
    grads, picks = [], []
    for impl, bptt in ((None, 1), ("ffma", 0)):
        m = M.CombinatorialRL(0, 256, K * N, 0, 10, 1, M.reward, "Dot", N, K, level="High" if high else "Low")
        m.load_state_dict(reference_shaped_state_dict(256, 8, 50 + it)); m = m.cuda().train()
        m.actor.impl = impl
        ops.set_option("bptt", bptt)
        R, ap, _, idx, _ = m(x, None, lat, sample="greedy", training="RL")
        w = torch.linspace(-0.5, 1.0, B, device="cuda")
        (w * sum(torch.log(p) for p in ap)).sum().backward()
        grads.append({k: p.grad.clone() for k, p in m.named_parameters()})
        picks.append(torch.stack(idx).clone())
    ops.set_option("bptt", 1)
    same_picks = bool(torch.equal(picks[0], picks[1]))
    worst = max(float((grads[0][k] - grads[1][k]).abs().max() / grads[1][k].abs().max().clamp(min=1e-3)) for k in grads[0])
    ok = worst < 1e-4 or not same_picks
    bad += not ok
    print(f"B={B} K={K} N={N} high={high}: picks equal {same_picks}, worst gradient deviation {worst:.2e} {'OK' if ok else 'VIOLATION'}", flush=True)
sys.exit(1 if bad else 0)
