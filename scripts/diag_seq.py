"""GPU-box bring-up diagnostic for the persistent tcgen05 scan kernels (not a test): compares the
persistent encoder / decoder (GNNPN_SEQ bits) against the strict-fp32 FFMA path on the same weights."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnnpn_sc_b200 import modelPN as M
from gnnpn_sc_b200.synth import pn_instances
from gnnpn_sc_b200.weights import reference_shaped_state_dict

for n, K, N in [(64, 6, 4), (300, 47, 5), (1000, 50, 10)]:
    x = pn_instances(n, K, N, seed=5).cuda()
    m = M.CombinatorialRL(0, 256, K * N, 0, 10, 1, M.reward, "Dot", N, K, level="Low")
    m.load_state_dict(reference_shaped_state_dict(256, 8, 77)); m = m.cuda().eval()
    outs = {}
    for impl in ("ffma", "tc"):
        m.actor.impl = impl
        with torch.no_grad():
            _, idx, lg = m.actor(x, None, sample="greedy")
        torch.cuda.synchronize()
        outs[impl] = (m.actor.last["enc_out"].clone(), torch.stack(idx), m.actor.last["win_logits"].clone(), m.actor.last["dec_h"].clone())
    e = (outs["tc"][0] - outs["ffma"][0]).abs()
    print(f"lstm n={n} K={K} N={N}: enc_out tc-vs-ffma max {e.max():.3e} (t=0 {e[:,0].max():.2e}, t=1 {e[:,1].max():.2e}, last {e[:,-1].max():.2e}); dec_h {(outs['tc'][3]-outs['ffma'][3]).abs().max():.3e}; "
          f"win_logits {(outs['tc'][2]-outs['ffma'][2]).abs().max():.3e}; picks differ {(outs['tc'][1]!=outs['ffma'][1]).sum().item()}", flush=True)
    if e.max() > 1e-3:
        bad = (e > 1e-3)
        print("   first bad t:", bad.any(2).any(0).nonzero().flatten()[:5].tolist(), "bad rows:", bad.any(2).any(1).nonzero().flatten()[:8].tolist(),
              "bad units:", bad.any(1).any(0).nonzero().flatten()[:16].tolist())
        print("   tc  [0,0,:8]", outs["tc"][0][0, 0, :8].tolist()); print("   ffma[0,0,:8]", outs["ffma"][0][0, 0, :8].tolist())
