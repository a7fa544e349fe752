"""Section timing of the differentiable replay (forward with saves, BPTT, weight-gradient contractions), B = 128, QWS."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnnpn_sc_b200 import modelPN as M, ops
from gnnpn_sc_b200.synth import pn_instances
from gnnpn_sc_b200.weights import reference_shaped_state_dict
K, N, H, F, B = 47, 5, 256, 8, int(sys.argv[1]) if len(sys.argv) > 1 else 128
L = K * N
dev = torch.device("cuda")
m = M.CombinatorialRL(0, H, L, 0, 10, 1, M.reward, "Dot", N, K, level="Low")
m.load_state_dict(reference_shaped_state_dict(H, F, 2)); m = m.to(dev).train()
a = m.actor
x = pn_instances(B, K, N, seed=5).to(dev)
idx = (torch.arange(K, device=dev).view(K, 1) * N + torch.randint(0, N, (K, B), device=dev)).to(torch.int32)
enc_w, dec_w = a._packed_weights()
def sync(): torch.cuda.synchronize(); return time.perf_counter()
for it in range(3):
    t0 = sync()
    sv = ops.pn_train_forward(x, enc_w, dec_w, idx, K, N)
    t1 = sync()
    gp = torch.rand(K, B, device=dev)
    dGe, dGd = ops.pn_train_backward(sv, gp, a.encoder.weight_hh_l0.detach(), a.decoder.weight_hh_l0.detach(), K, N)
    t2 = sync()
    enc_lnh = sv["enc_out"].permute(1, 0, 2)
    h_prev_e = torch.cat([torch.zeros_like(enc_lnh[:1]), enc_lnh[:-1]]).reshape(L * B, H).t().contiguous()
    x_e = x.permute(1, 0, 2).reshape(L * B, F).t().contiguous()
    t3 = sync()
    dWhh = ops.gemm_bias_act(dGe, h_prev_e, impl="tc")
    t4 = sync()
    dM = ops.gemm_bias_act(dGe, x_e, impl="ffma")
    t5 = sync()
    db = dGe.sum(1)
    t6 = sync()
    print(f"B={B}: forward+saves {1e3*(t1-t0):.2f} ms | BPTT scans {1e3*(t2-t1):.2f} | transposes {1e3*(t3-t2):.2f} | dW_hh gemm(tc) {1e3*(t4-t3):.2f} | "
          f"dM gemm(ffma) {1e3*(t5-t4):.2f} | bias sum {1e3*(t6-t5):.2f}", flush=True)
