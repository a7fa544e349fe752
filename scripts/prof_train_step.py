"""Where one REINFORCE step spends its time (wall, with synchronisation between the sections)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnnpn_sc_b200 import modelPN as M, ops
from gnnpn_sc_b200.synth import pn_instances
from gnnpn_sc_b200.weights import reference_shaped_state_dict
K, N, H, F, B = 47, 5, 256, 8, 128
dev = torch.device("cuda")
m = M.CombinatorialRL(0, H, K * N, 0, 10, 1, M.reward, "Dot", N, K, level="Low")
m.load_state_dict(reference_shaped_state_dict(H, F, 2)); m = m.to(dev).train()
m.actor.replay_impl = sys.argv[1] if len(sys.argv) > 1 else "own"
m.actor.check_inputs = False
opt = torch.optim.Adam(m.actor.parameters(), lr=1e-4)
x = pn_instances(B, K, N, seed=5).to(dev)
def sync(): torch.cuda.synchronize(); return time.perf_counter()
for it in range(4):
    t0 = sync()
    with torch.no_grad():
        probs, idx, logits = m.actor(x, None, sample="sample")
    t1 = sync()
    idxs = torch.stack(idx)
    ap = m.actor.replay_action_probs(x, idxs, None)
    t2 = sync()
    rows = torch.arange(B, device=dev)
    actions = list(x[rows.unsqueeze(0), idxs].unbind(0))
    R = m.reward(actions, None, K, USE_CUDA=True, level="Low", embedding_size=0)
    t3 = sync()
    logp = 0
    for p in ap: logp = logp + torch.log(p)
    loss = ((R - R.mean()) * logp).mean()
    opt.zero_grad(); loss.backward()
    t4 = sync()
    torch.nn.utils.clip_grad_norm_(m.actor.parameters(), 2.0); opt.step()
    t5 = sync()
    print(f"[{m.actor.replay_impl}] sampled decode {1e3*(t1-t0):.2f} ms | replay forward {1e3*(t2-t1):.2f} | reward {1e3*(t3-t2):.2f} | "
          f"loss+backward {1e3*(t4-t3):.2f} | clip+adam {1e3*(t5-t4):.2f}", flush=True)
