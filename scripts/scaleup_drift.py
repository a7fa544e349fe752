"""Rounding drift of the scale-up configuration (K = 100 tasks x N = 1000 candidates, L = 100,000 encoder steps) against
float64 (SURVEY 7 hard part 4): the encoder's hidden states after 10^2 .. 10^5 dependent steps, and the first decode
step's 1000 window logits, for a few instances -- GPU (tcgen05 column-split scan, 3xFP16 split, MUFU activations) and
torch fp32 on the CPU, both against a float64 evaluation of the same LSTM (oracle.pn_oracle.lstm_cell_f64; the CPU legs
are the checker, not the product).

    python scripts/scaleup_drift.py [--out gpurun_out/scaleup_drift.json]"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnnpn_sc_b200 import modelPN as M
from gnnpn_sc_b200.synth import pn_instances
from oracle import pn_oracle as po

ap = argparse.ArgumentParser()
ap.add_argument("--out", default="gpurun_out/scaleup_drift.json")
ap.add_argument("--instances", type=int, default=4)
a = ap.parse_args()
K, N, n = 100, 1000, a.instances
L = K * N
cfg = po.PNConfig(seq_len=L, s_number=N, s_category=K)
sd = po.make_state_dict(cfg, 1)
x = pn_instances(n, K, N, seed=77)
m = M.CombinatorialRL(0, 256, L, 0, 10, 1, M.reward, "Dot", N, K, level="Low")
m.load_state_dict(sd)
m = m.cuda().eval()
with torch.no_grad():
    m(x.cuda(), None, sample="greedy", training="SL")
last = m.actor.last
enc_gpu = last["enc_out"].cpu()
wl_gpu = last["win_logits"].cpu()[:, :N]
q_gpu = last["dec_h"].cpu()[:, 0]

# float64 and float32 CPU evaluations of the encoder (embedding2 + LSTM), step by step
emb64 = po.embed_inputs({k: v.double() for k, v in sd.items()}, cfg, x.double())
w = [sd[f"actor.encoder.{k}"].double() for k in ("weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0")]
checkpoints = [99, 999, 9999, 99999]
h64 = torch.zeros(n, 256, dtype=torch.float64); c64 = h64.clone()
t0 = time.time()
snap64 = {}
for t in range(L):
    h64, c64 = po.lstm_cell_f64(*w, emb64[:, t], h64, c64)
    if t in checkpoints:
        snap64[t] = h64.clone()
with torch.no_grad():
    enc32, _ = po._lstm(sd, "encoder", po.embed_inputs(sd, cfg, x))
res = {"config": {"K": K, "N": N, "L": L, "instances": n}, "encoder_hidden_state_max_abs_error_vs_f64": {},
       "cpu_seconds_f64": time.time() - t0}
for t in checkpoints:
    res["encoder_hidden_state_max_abs_error_vs_f64"][f"t={t + 1}"] = {
        "gpu_tcgen05": float((enc_gpu[:, t].double() - snap64[t]).abs().max()),
        "torch_cpu_fp32": float((enc32[:, t].double() - snap64[t]).abs().max())}
# first decode step in float64 from the float64 encoder state: query and the 1000 window logits
wd = [sd[f"actor.decoder.{k}"].double() for k in ("weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0")]
start = sd["actor.decoder_start_input"].double().unsqueeze(0).expand(n, -1)
q64, _ = po.lstm_cell_f64(*wd, start, h64, c64)
# window rows of step 0 in float64 need the first N encoder states
h, c = torch.zeros(n, 256, dtype=torch.float64), torch.zeros(n, 256, dtype=torch.float64)
rows64 = []
for t in range(N):
    h, c = po.lstm_cell_f64(*w, emb64[:, t], h, c)
    rows64.append(h.clone())
rows64 = torch.stack(rows64, 1)                                   # [n, N, H]
logits64 = 10.0 * torch.tanh(torch.einsum("bnh,bh->bn", rows64, q64))
res["first_decode_step"] = {
    "query_max_abs_error_vs_f64": float((q_gpu.double() - q64).abs().max()),
    "window_logits_max_abs_error_vs_f64": float((wl_gpu.double() - logits64).abs().max()),
    "window_logits_max_rel_error_vs_f64": float(((wl_gpu.double() - logits64).abs() / logits64.abs().clamp(min=1)).max()),
    "picks_equal_f64_argmax": bool((wl_gpu.argmax(1) == logits64.argmax(1)).all())}
print(json.dumps(res, indent=1))
os.makedirs(os.path.dirname(a.out), exist_ok=True)
with open(a.out, "w") as f:
    json.dump(res, f, indent=1)
