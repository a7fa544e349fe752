"""GPU bring-up diagnostic for the column-split cluster scan (tc_colsplit.cu): encoder outputs against the
CTA-pair persistent scan (GNNPN_COLSPLIT=0) and the strict-fp32 FFMA path, then timing of both over batch sizes."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnnpn_sc_b200 import ops, modelPN as M
from gnnpn_sc_b200.synth import pn_instances
from gnnpn_sc_b200.weights import reference_shaped_state_dict

dev = torch.device("cuda")
m = M.CombinatorialRL(0, 256, 235, 0, 10, 1, M.reward, "Dot", 5, 47, level="Low")
m.load_state_dict(reference_shaped_state_dict(256, 8, 77)); m = m.cuda().eval()
enc_w, dec_w = m.actor._packed_weights()


def encode(x, mode, ws, g=None):
    os.environ["GNNPN_COLSPLIT"] = str(mode)
    if g is None:
        os.environ.pop("GNNPN_COLSPLIT_G", None)
    else:
        os.environ["GNNPN_COLSPLIT_G"] = str(g)
    out = ops.lstm_encode(x, enc_w, 256, workspace=ws)
    torch.cuda.synchronize()
    return out


if "--time-only" not in sys.argv:
    for n, K, N in [(1, 3, 2), (128, 6, 4), (300, 47, 5), (1000, 20, 5), (2100, 9, 5)]:
        x = pn_instances(n, K, N, seed=5).to(dev)
        ws = ops.pn_workspace(n, 256, dev, "tc")
        e_g2, c_g2 = encode(x, 1, ws, 2)
        e_g2, c_g2 = e_g2.clone(), c_g2.clone()
        e_cs, c_cs = encode(x, 1, ws, 1)
        e_cs, c_cs = e_cs.clone(), c_cs.clone()
        e_sq, c_sq = encode(x, 0, ws)
        e_ff, c_ff = ops.lstm_encode(x, enc_w, 256, workspace=None)
        torch.cuda.synchronize()
        d1 = (e_cs - e_sq).abs(); d2 = (e_cs - e_ff).abs(); d3 = (e_sq - e_ff).abs()
        print(f"n={n} L={K*N}: two-group clusters vs pair: enc {(e_g2 - e_sq).abs().max():.2e} c {(c_g2 - c_sq).abs().max():.2e}", flush=True)
        print(f"n={n} L={K*N}: colsplit-vs-pair enc {d1.max():.2e} c {(c_cs-c_sq).abs().max():.2e} | colsplit-vs-ffma {d2.max():.2e} "
              f"(t=0 {d2[:,0].max():.1e}, t=1 {d2[:,1].max():.1e}, last {d2[:,-1].max():.1e}) | pair-vs-ffma {d3.max():.2e}", flush=True)
        if d2.max() > 1e-3:
            bad = d2 > 1e-3
            print("   first bad t:", bad.any(2).any(0).nonzero().flatten()[:5].tolist(), "rows:", bad.any(2).any(1).nonzero().flatten()[:8].tolist(),
                  "units:", bad.any(1).any(0).nonzero().flatten()[:16].tolist())

if "--time-only" not in sys.argv:
    # full greedy forward (encoder + fused decode) through the module: column-split vs CTA-pair kernels
    for n, K, N in [(1, 3, 2), (128, 6, 4), (300, 47, 5), (1000, 50, 10), (130, 12, 32)]:
        x = pn_instances(n, K, N, seed=7).to(dev)
        mm = M.CombinatorialRL(0, 256, K * N, 0, 10, 1, M.reward, "Dot", N, K, level="High")
        mm.load_state_dict(reference_shaped_state_dict(256, 8, 78)); mm = mm.cuda().eval()
        lat = [torch.randn(n, K * N, device=dev) for _ in range(K)]
        outs = {}
        for mode in (1, 0):
            os.environ["GNNPN_COLSPLIT"] = str(mode)
            with torch.no_grad():
                _, idx, lg = mm.actor(x, lat, sample="greedy")
            torch.cuda.synchronize()
            outs[mode] = (torch.stack(idx).clone(), mm.actor.last["win_logits"].clone(), mm.actor.last["dec_h"].clone(),
                          mm.actor.last["win_probs"].clone())
        print(f"decode n={n} K={K} N={N}: picks differ {(outs[1][0] != outs[0][0]).sum().item()}/{outs[0][0].numel()}, "
              f"dec_h {(outs[1][2]-outs[0][2]).abs().max():.2e}, win_logits {(outs[1][1]-outs[0][1]).abs().max():.2e}, "
              f"win_probs {(outs[1][3]-outs[0][3]).abs().max():.2e}", flush=True)

rows = []
for n in [128, 1920, 2048, 3840, 4096, 7680, 8192]:
    x = pn_instances(n, 47, 5, seed=5).to(dev)
    ws = ops.pn_workspace(n, 256, dev, "tc")
    enc_out = torch.empty(n, 235, 256, device=dev); c = torch.empty(n, 256, device=dev)
    r = {"n": n, "L": 235}
    for mode, name, g in ((1, "colsplit_ms", 1), (1, "colsplit_g2_ms", 2), (0, "pair_ms", None)):
        os.environ["GNNPN_COLSPLIT"] = str(mode)
        if g is None:
            os.environ.pop("GNNPN_COLSPLIT_G", None)
        else:
            os.environ["GNNPN_COLSPLIT_G"] = str(g)
        ts = []
        for i in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); ops.lstm_encode(x, enc_w, 256, enc_out, c, workspace=ws); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        r[name] = sorted(ts[1:])[len(ts[1:]) // 2]
    os.environ.pop("GNNPN_COLSPLIT_G", None)
    r["us_per_step_colsplit"] = r["colsplit_ms"] * 1e3 / 235
    r["us_per_step_pair"] = r["pair_ms"] * 1e3 / 235
    # fused greedy decode (K = 47 steps incl. the pointer phase)
    dec_h = torch.empty(n, 47, 256, device=dev); idx = torch.empty(47, n, device=dev, dtype=torch.int32)
    wl = torch.empty(n, 235, device=dev); wp = torch.empty(n, 235, device=dev)
    for mode, name in ((1, "dec_colsplit_ms"), (0, "dec_pair_ms")):
        os.environ["GNNPN_COLSPLIT"] = str(mode)
        ts = []
        for i in range(5):
            ops.lstm_encode(x, enc_w, 256, enc_out, c, workspace=ws)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); ops.pn_decode_greedy(x, enc_out, c, dec_w, 47, 5, out=(dec_h, idx, wl, wp), workspace=ws); b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        r[name] = sorted(ts[1:])[len(ts[1:]) // 2]
    rows.append(r)
    print(json.dumps(r), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/colsplit_timing.jsonl", "w") as f:
    for r in rows:
        f.write(json.dumps(r) + "\n")
