"""torchrun --nproc-per-node 2 scripts/dp_check.py : data-parallel REINFORCE step on 2 GPUs == the single-GPU
step on the concatenated batch (gradients after the bucket all-reduce agree to fp32 reduction-order noise, EMA-baseline input), and instance-sharded
greedy decode == unsharded decode (bit-identical picks)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from gnnpn_sc_b200 import modelPN as M, parallel
from gnnpn_sc_b200.synth import pn_instances
from gnnpn_sc_b200.weights import reference_shaped_state_dict

def build(K, N):
    m = M.CombinatorialRL(0, 256, K * N, 0, 10, 1, M.reward, "Dot", N, K, level="High")
    m.load_state_dict(reference_shaped_state_dict(256, 8, 3))
    return m.cuda()

def step(m, x):
    m.train(); m.zero_grad()
    R, probs, actions, idx, _ = m(x, None, None, sample="greedy")      # deterministic picks, differentiable probs
    r_mean = parallel.global_mean(R)
    logp = sum(torch.log(p) for p in probs)
    ((R - r_mean) * logp).mean().backward()
    parallel.allreduce_gradients(m.actor.parameters())
    return torch.cat([p.grad.reshape(-1) for p in m.actor.parameters()]), r_mean, torch.stack(idx)

local = int(os.environ.get("LOCAL_RANK", 0)); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
K, N, B = 12, 5, 256
x = pn_instances(B, K, N, seed=8).cuda()
m = build(K, N)
g_dp, r_dp, idx_dp = step(m, parallel.shard(x))
idx_all = parallel.gather_concat(idx_dp.t().contiguous()).t()
if dist.get_rank() == 0:
    import types
    saved = parallel._on
    parallel._on = lambda: False                     # single-process reference on the full batch
    g_1, r_1, idx_1 = step(build(K, N), x)
    parallel._on = saved
    rel = ((g_dp - g_1).abs().max() / g_1.abs().max()).item()
    print(f"DP(2) vs single: grad max rel dev {rel:.2e}, reward mean {r_dp.item():.6f} vs {r_1.item():.6f}, "
          f"picks identical: {bool(torch.equal(idx_all, idx_1))}")
    assert rel < 5e-5 and abs(r_dp.item() - r_1.item()) < 1e-6 and torch.equal(idx_all, idx_1)
    print("DP_CHECK_OK")
dist.barrier(); dist.destroy_process_group()
