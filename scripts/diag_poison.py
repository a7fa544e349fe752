"""Does any full-wave launch read memory it has not written?  Reference results are computed in a clean process state, then
the caching allocator's free blocks are filled with a poison pattern and the blocked encoder + fused decoder (PNLow and
PNHigh with latent) run on buffers carved from the poisoned blocks; outputs are compared bitwise."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnnpn_sc_b200 import ops, modelPN as M
from gnnpn_sc_b200.synth import pn_instances
from gnnpn_sc_b200.weights import reference_shaped_state_dict

dev = torch.device("cuda")
K, N = 47, 5
n = int(os.environ.get("DIAG_N", "18944"))
L = K * N
x = pn_instances(n, K, N, seed=1234).to(dev)
nets = []
for level, seed in (("Low", 1), ("High", 2)):
    m = M.CombinatorialRL(0, 256, L, 0, 10, 1, M.reward, "Dot", N, K, level=level)
    m.load_state_dict(reference_shaped_state_dict(256, 8, seed))
    nets.append(m.cuda().eval())


def run():
    with torch.no_grad():
        _, _, _, i_lo, lat = nets[0](x, None, sample="greedy", training="SL")
        R, _, _, i_hi, _ = nets[1](x, None, lat, sample="greedy", training="RL")
    torch.cuda.synchronize()
    out = {"idx_lo": torch.stack(i_lo).clone(), "idx_hi": torch.stack(i_hi).clone(), "R": R.clone()}
    for tag, m in (("lo", nets[0]), ("hi", nets[1])):
        out["wl_" + tag] = m.actor.last["win_logits"].clone()
        out["wp_" + tag] = m.actor.last["win_probs"].clone()
        m.actor.last = None
    return out


ref = run()
for name, fill in (("nan", float("nan")), ("1e30", 1e30), ("one", 1.0), ("randn", None), ("zero", 0.0)):
    torch.cuda.empty_cache()
    junk = [torch.empty(1 << 28, device=dev) for _ in range(24)]          # 24 GiB of 1 GiB blocks
    junk += [torch.empty(1 << 20, device=dev) for _ in range(256)]        # and small ones
    for j in junk:
        j.normal_() if fill is None else j.fill_(fill)
    torch.cuda.synchronize()
    del junk                                                              # back to the cache, contents intact
    got = run()
    msg = []
    for k in ref:
        a, b = got[k].float(), ref[k].float()
        bad = (a != b) & ~(torch.isnan(a) & torch.isnan(b))
        if bad.any():
            nz = bad.nonzero()
            msg.append(f"{k}: {int(bad.sum())} differ (max |d| {(a.double() - b.double()).abs()[bad].max().item():.3e}; first at {nz[0].tolist()}, "
                       f"instances%128 {sorted(set((nz[:, 0 if a.shape[0] == n else -1] % 128).tolist()))[:10]})")
    print(f"poison {name}: " + ("all outputs bit-identical to the clean run" if not msg else " | ".join(msg)), flush=True)
