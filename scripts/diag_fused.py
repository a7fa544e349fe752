"""GPU bring-up diagnostic for the fused CTA-pair decoder (blocked encodings, pointer dots in the cell epilogue,
tc_seq.cu): encodings / decoder states / window logits / picks against the column-split cluster scan (bitwise) and the
strict-fp32 FFMA path, then launch times of the encoder and the fused decoder at a full wave.

    python scripts/diag_fused.py [--time-only] [--prof]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnnpn_sc_b200 import ops, modelPN as M
from gnnpn_sc_b200.synth import pn_instances
from gnnpn_sc_b200.weights import reference_shaped_state_dict

dev = torch.device("cuda")


def run(mm, x, lat, scan, impl="tc", sample="greedy", forced=None):
    ops.set_option("scan", scan)
    mm.actor.impl = impl
    with torch.no_grad():
        _, idx, _ = mm.actor(x, lat, sample=sample, forced_idxs=forced)
    torch.cuda.synchronize()
    last = mm.actor.last
    return {"idx": torch.stack(idx).clone(), "enc_out": last["enc_out"].clone(), "dec_h": last["dec_h"].clone(),
            "win_logits": last["win_logits"].clone(), "win_probs": last["win_probs"].clone(), "layout": last["enc_layout"]}


if "--time-only" not in sys.argv:
    for n, K, N in [(1, 3, 2), (128, 6, 4), (300, 47, 5), (257, 50, 10), (130, 12, 8), (129, 9, 7), (200, 5, 12)]:
        x = pn_instances(n, K, N, seed=7).to(dev)
        mm = M.CombinatorialRL(0, 256, K * N, 0, 10, 1, M.reward, "Dot", N, K, level="High")
        mm.load_state_dict(reference_shaped_state_dict(256, 8, 78)); mm = mm.cuda().eval()
        lat = [torch.randn(n, K * N, device=dev) for _ in range(K)]
        pair, cs, ff = run(mm, x, lat, 0), run(mm, x, lat, 1), run(mm, x, lat, -1, impl="ffma")
        msg = [f"n={n} K={K} N={N} layout(pair)={pair['layout']}"]
        for k in ("idx", "enc_out", "dec_h", "win_logits", "win_probs"):
            a, b, c = pair[k].float(), cs[k].float(), ff[k].float()
            msg.append(f"{k}: pair-vs-cs {(a - b).abs().max().item():.2e} pair-vs-ffma {(a - c).abs().max().item():.2e}")
        print(" | ".join(msg), flush=True)
        bad = (pair["win_logits"] != cs["win_logits"])
        if bad.any():
            rows = bad.any(1).nonzero().flatten()[:8].tolist()
            cols = bad.any(0).nonzero().flatten()[:8].tolist()
            print("   win_logits differ at rows", rows, "cols", cols, flush=True)
    ops.set_option("scan", -1)

K, N, n = 47, 5, 18944
if "--normal" in sys.argv:
    K, N = 50, 10
x = pn_instances(n, K, N, seed=5).to(dev)
mm = M.CombinatorialRL(0, 256, K * N, 0, 10, 1, M.reward, "Dot", N, K, level="Low")
mm.load_state_dict(reference_shaped_state_dict(256, 8, 77)); mm = mm.cuda().eval()
enc_w, dec_w = mm.actor._packed_weights()
ws = ops.pn_workspace(n, 256, dev, "tc")
c = torch.empty(n, 256, device=dev)
out = (torch.empty(n, K, 256, device=dev), torch.empty(K, n, device=dev, dtype=torch.int32),
       torch.empty(n, K * N, device=dev), torch.empty(n, K * N, device=dev))
variants = [(ops.ENC_BLOCKED128, "blocked/fused, dec_h not stored", False), (ops.ENC_BLOCKED128, "blocked/fused, dec_h stored", True),
            (ops.ENC_ROWMAJOR, "row-major/pointer phase", True)]
for lay, name, with_h in variants:
    ops.set_option("scan", 0)
    out = (torch.empty(n, K, 256, device=dev) if with_h else None,) + tuple(out[1:])
    enc = ops.enc_out_empty(n, K * N, 256, lay, dev)
    te, td = [], []
    for i in range(6):
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record(); ops.lstm_encode(x, enc_w, 256, enc, c, workspace=ws, layout=lay); e1.record()
        ops.pn_decode_greedy(x, enc, c, dec_w, K, N, out=out, workspace=ws, enc_layout=lay); e2.record()
        torch.cuda.synchronize()
        if i >= 2:
            te.append(e0.elapsed_time(e1)); td.append(e1.elapsed_time(e2))
    print(f"n={n} K={K} N={N} {name}: encoder {min(te):.3f} ms, decoder {min(td):.3f} ms", flush=True)
    if "--prof" in sys.argv:
        ops.set_option("prof", 1)
        ops.lstm_encode(x, enc_w, 256, enc, c, workspace=ws, layout=lay)
        ops.pn_decode_greedy(x, enc, c, dec_w, K, N, out=out, workspace=ws, enc_layout=lay)
        torch.cuda.synchronize()
        ops.set_option("prof", 0)
    del enc
ops.set_option("scan", -1)
