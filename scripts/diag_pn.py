"""GPU-box diagnostic: error statistics of the CUDA PN path against the CPU oracle (not a test)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import pn_oracle as po
from gnnpn_sc_b200 import modelPN as M
from gnnpn_sc_b200.synth import pn_instances

def tf(n, K, N, gain):
    cfg = po.PNConfig(seq_len=K * N, s_number=N, s_category=K)
    sd = po.make_state_dict(cfg, 77, gain)
    x = pn_instances(n, K, N, seed=5)
    with torch.no_grad():
        _, idx_ref, lg_ref, internals = po.pointer_forward(sd, cfg, x, None, "greedy", return_internals=True)
    m = M.CombinatorialRL(0, 256, K * N, 0, 10, 1, M.reward, "Dot", N, K, level="Low")
    m.load_state_dict(sd); m = m.cuda().eval()
    with torch.no_grad():
        probs, idx, lg = m.actor(x.cuda(), None, sample="greedy", forced_idxs=[t.cuda() for t in idx_ref])
    enc = m.actor.last["enc_out"].cpu()
    e_enc = (enc - internals["enc_out"]).abs()
    print(f"[n={n} K={K} N={N} gain={gain}] enc_out max abs err {e_enc.max():.3e} (per t: first {e_enc[:,0].max():.2e} last {e_enc[:,-1].max():.2e})")
    q = torch.stack(internals["queries"], 1)
    e_q = (m.actor.last["dec_h"].cpu() - q).abs()
    print(f"   dec_h max abs err {e_q.max():.3e}")
    dense = torch.stack([lg[k] for k in range(K)]).cpu().numpy(); ref = torch.stack(lg_ref).numpy()
    fin = np.isfinite(ref)
    err = np.abs(dense[fin] - ref[fin]); rel = err / np.maximum(1, np.abs(ref[fin]))
    print(f"   logits: max abs {err.max():.3e} max rel {rel.max():.3e}  frac>1e-5: {(rel>1e-5).mean():.2e}  |ref| max {np.abs(ref[fin]).max():.3f}")
    # pre-tanh dot product error
    u_ref = torch.einsum('blh,bkh->kbl', internals["enc_out"].double(), q.double()).numpy()
    worst = np.argmax(np.where(fin, np.abs(dense - ref), 0)); k, b, l = np.unravel_index(worst, ref.shape)
    print(f"   worst at k={k} b={b} l={l}: mine {dense[k,b,l]:.7f} ref {ref[k,b,l]:.7f} u(exact from ref enc) {u_ref[k,b,l]:.6f}")
    i_m = torch.stack(idx).cpu().numpy(); i_r = torch.stack(idx_ref).numpy()
    print(f"   picks differing: {(i_m != i_r).sum()} / {i_r.size}")
    # fp64 truth for both
    with torch.no_grad():
        e = po.embed_inputs(sd, cfg, x).double()
        h = torch.zeros(n, 256, dtype=torch.float64); c = torch.zeros(n, 256, dtype=torch.float64)
        outs = []
        for t in range(K * N):
            h, c = po.lstm_cell_f64(sd["actor.encoder.weight_ih_l0"], sd["actor.encoder.weight_hh_l0"], sd["actor.encoder.bias_ih_l0"], sd["actor.encoder.bias_hh_l0"], e[:, t], h, c)
            outs.append(h)
        truth = torch.stack(outs, 1)
    print(f"   vs fp64 truth: cuda enc err {(enc.double()-truth).abs().max():.3e}, torch-cpu enc err {(internals['enc_out'].double()-truth).abs().max():.3e}")

def rew():
    n, K, N = 512, 47, 5
    x = pn_instances(n, K, N, seed=9)
    g = torch.Generator().manual_seed(0)
    idx = (torch.arange(K).view(K, 1) * N + torch.randint(0, N, (K, n), generator=g)).int()
    from gnnpn_sc_b200 import ops
    viol, obj, r = ops.pn_reward(x.cuda(), idx.cuda())
    acts = [x[torch.arange(n), idx[k].long()] for k in range(K)]
    v_ref, o_ref = po.composition_objective(torch.stack(acts).numpy())
    r_ref = po.reward(acts, None, K, "High", 0)
    print("reward: viol equal", np.array_equal(viol.cpu().numpy(), v_ref), "obj equal", np.array_equal(obj.cpu().numpy(), o_ref), "rew equal", torch.equal(r.cpu(), r_ref))
    bad = np.nonzero(obj.cpu().numpy() != o_ref)[0]
    for b in bad[:5]:
        q0 = torch.stack(acts)[:, b, 0].numpy()
        print("  b", b, "mine", obj[b].item(), "ref", o_ref[b], "np.sum", np.sum(q0), "seq", float(np.cumsum(q0, dtype=np.float32)[-1]), "used", (q0 > 0).sum())
    bad = np.nonzero(r.cpu().numpy() != r_ref.numpy())[0]
    for b in bad[:5]:
        print("  rew b", b, "mine", repr(r[b].item()), "ref", repr(r_ref[b].item()), "viol", v_ref[b], "obj", repr(float(o_ref[b])))

rew()
tf(129, 50, 10, 1.0)
tf(64, 12, 5, 3.0)
tf(300, 47, 5, 1.0)
