"""REINFORCE training steps / s (trainPNLow.py:77-106 / trainPNHigh.py:77-112) on synthetic QWS-shaped batches, B = 128 per
rank (the reference's batch size): sampled decode (tcgen05 kernels) -> reward kernel -> differentiable replay + BPTT ->
gradient all-reduce (NCCL, when launched under torchrun) -> clip -> Adam.

    python scripts/bench_train.py [--impl own|torch] [--level low|high] [--steps 20]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_train.py ...

--impl own: the library's forward-with-saves + BPTT kernels (gnnpn_pn_train_*); torch: nn.LSTM (cuDNN) + torch ops replay."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from gnnpn_sc_b200 import modelPN as M, parallel, trainPN, _lib
from gnnpn_sc_b200.synth import pn_instances
from gnnpn_sc_b200.weights import reference_shaped_state_dict

ap = argparse.ArgumentParser()
ap.add_argument("--impl", default="own", choices=["own", "torch"])
ap.add_argument("--level", default="low", choices=["low", "high"])
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--batch", type=int, default=128)
ap.add_argument("--out", default=None)
a = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
K, N, H, F = 47, 5, 256, 8


def net(level, seed):
    m = M.CombinatorialRL(0, H, K * N, 0, 10, 1, M.reward, "Dot", N, K, level=level)
    m.load_state_dict(reference_shaped_state_dict(H, F, seed))
    m.actor.replay_impl = a.impl
    m.actor.check_inputs = False
    return m.to(dev)


model = net("High" if a.level == "high" else "Low", 2)
low = net("Low", 1).eval() if a.level == "high" else None
tm = trainPN.TrainModel(model, [0], [0], 1, 0.9, True, "bench", K, lr=1e-4, batch_size=a.batch, low_model=low, root="/tmp")
x = pn_instances(a.batch, K, N, seed=5 + rank).to(dev)
baseline = torch.zeros(1, device=dev)
model.train()
for i in range(3):
    _, baseline = tm.reinforce_step(x, None, baseline, i == 0)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
l0 = _lib.launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for i in range(a.steps):
    r, baseline = tm.reinforce_step(x, None, baseline, False)
e1.record()
torch.cuda.synchronize()
wall = time.perf_counter() - t0
ms = torch.tensor([e0.elapsed_time(e1) / a.steps], device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    line = {"workload": f"reinforce_step_pn{a.level}_qws", "impl": a.impl, "batch_per_rank": a.batch, "n_gpus": world,
            "ms_per_step": float(ms.item()), "steps_per_s": 1e3 / float(ms.item()), "instances_per_s": world * a.batch * 1e3 / float(ms.item()),
            "wall_ms_per_step": wall * 1e3 / a.steps, "library_launches_per_step": (_lib.launch_count() - l0) / a.steps,
            "mean_reward": float(r)}
    print(json.dumps(line))
    if a.out:
        with open(a.out, "a") as f:
            f.write(json.dumps(line) + "\n")
if world > 1:
    dist.destroy_process_group()
