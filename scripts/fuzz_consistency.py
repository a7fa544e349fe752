"""Randomised cross-kernel consistency: for random (n, K, N) the tensor-core dispatch (auto), the CTA-pair scan (scan = 0) and
the strict-fp32 kernels must give bit-identical results among the tensor-core scans and <= 1e-5 against the FFMA path
(picks may differ only on sub-tolerance margins).  Prints one line per configuration; exits non-zero on a violation."""
import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnnpn_sc_b200 import modelPN as M, ops
from gnnpn_sc_b200.synth import pn_instances
from gnnpn_sc_b200.weights import reference_shaped_state_dict

rng = random.Random(int(os.environ.get("FUZZ_SEED", "1")))
bad = 0
for it in range(int(os.environ.get("FUZZ_ITERS", "14"))):
    n = rng.choice([1, 2, 31, 127, 128, 129, 255, 300, 517, 1000, 1300, 2049])
    K = rng.randint(1, 12)
    N = rng.choice([1, 2, 3, 5, 7, 8, 9, 10, 11, 16, 31, 32])
    x = pn_instances(n, K, N, seed=it).cuda()
    m = M.CombinatorialRL(0, 256, K * N, 0, 10, 1, M.reward, "Dot", N, K, level="High")
    m.load_state_dict(reference_shaped_state_dict(256, 8, 100 + it)); m = m.cuda().eval()
    lat = [torch.randn(n, K * N, device="cuda") for _ in range(K)]
    outs = {}
    for key, (scan, impl) in {"auto": (-1, None), "pair": (0, None), "ffma": (-1, "ffma")}.items():
        ops.set_option("scan", scan); m.actor.impl = impl
        with torch.no_grad():
            _, idx, _ = m.actor(x, lat, sample="greedy")
        last = m.actor.last
        outs[key] = (torch.stack(idx).clone(), last["win_logits"].clone(), last["win_probs"].clone(), last["enc_out"].clone(), last["dec_h"].clone())
    ops.set_option("scan", -1)
    same = all(torch.equal(a, b) for a, b in zip(outs["auto"], outs["pair"]))
    d_log = float((outs["auto"][1] - outs["ffma"][1]).abs().max())
    d_enc = float((outs["auto"][3] - outs["ffma"][3]).abs().max())
    flips = int((outs["auto"][0] != outs["ffma"][0]).sum())
    ok = same and d_log <= 1e-4 and d_enc <= 1e-5
    bad += not ok
    print(f"n={n} K={K} N={N}: tc scans bit-identical {same}; vs ffma logits {d_log:.2e} enc {d_enc:.2e} pick diffs {flips} {'OK' if ok else 'VIOLATION'}", flush=True)
sys.exit(1 if bad else 0)
