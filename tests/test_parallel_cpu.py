"""world_size-2 gloo tests (CPU) of the multi-GPU plumbing: instance sharding and the data-parallel
REINFORCE synchronisation (flat-bucket gradient all-reduce + global reward mean)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gnnpn_sc_b200 import parallel


def test_shard_ranges_partition_the_batch():
    for n in (0, 1, 7, 128, 129, 18944):
        for w in (1, 2, 4, 8):
            spans = [parallel.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(hi - lo for lo, hi in spans) <= -(-n // w) if n else True


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 1))
    g = torch.Generator().manual_seed(1)
    x = torch.randn(10, 6, generator=g)                     # ragged: 5 + 5 at w=2, uneven below
    R = torch.randn(10, generator=g)
    xs, Rs = parallel.shard(x), parallel.shard(R)
    r_mean = parallel.global_mean(Rs)
    loss = ((Rs - r_mean) * model(xs).squeeze(1)).mean()
    loss.backward()
    parallel.allreduce_gradients(model.parameters())
    flat = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    gathered = parallel.gather_concat(xs[:, 0].contiguous())
    if rank == 0:
        torch.save({"grad": flat, "r_mean": r_mean, "gathered": gathered}, out)
    dist.destroy_process_group()


def test_data_parallel_step_equals_single_process(tmp_path):
    out = str(tmp_path / "dp.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 1))
    g = torch.Generator().manual_seed(1)
    x = torch.randn(10, 6, generator=g)
    R = torch.randn(10, generator=g)
    loss = ((R - R.mean()) * model(x).squeeze(1)).mean()
    loss.backward()
    want = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    assert torch.allclose(got["r_mean"], R.mean(), atol=1e-7)
    assert torch.allclose(got["grad"], want, atol=1e-6)      # equal shards: mean of shard means == global mean
    assert torch.equal(got["gathered"], x[:, 0])


def _ragged_worker(rank, world, port, out, n):
    """The trainer's recipe (trainPN.TrainModel.reinforce_step): different initial weights and RNG state per rank,
    a batch of ``n`` rows that does not divide by the world size (n = 1: rank 1's shard is EMPTY)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(100 + rank)                           # replicas start DIFFERENT ...
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 1))
    parallel.broadcast_parameters(model.parameters())       # ... and are made identical
    seed = parallel.shared_seed()
    g = torch.Generator().manual_seed(1)
    x = torch.randn(n, 6, generator=g)
    R = torch.randn(n, generator=g)
    xs, Rs = parallel.shard(x), parallel.shard(R)
    r_mean, n_global = parallel.global_mean_count(Rs)
    if xs.shape[0]:
        loss = ((Rs - r_mean) * model(xs).squeeze(1)).sum() / n_global
        loss.backward()
    parallel.allreduce_gradients(list(model.parameters()), average=False)
    flat = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    w0 = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    torch.save({"grad": flat, "r_mean": r_mean, "w0": w0, "seed": seed, "n_local": xs.shape[0]}, f"{out}.{rank}")
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [7, 1])
def test_ragged_and_empty_shards_equal_single_process(tmp_path, n):
    out = str(tmp_path / "ragged.pt")
    mp.spawn(_ragged_worker, args=(2, _free_port(), out, n), nprocs=2, join=True)
    a, b = torch.load(out + ".0"), torch.load(out + ".1")
    assert a["n_local"] + b["n_local"] == n and (n != 1 or b["n_local"] == 0)
    assert torch.equal(a["w0"], b["w0"])                    # broadcast: replicas identical
    assert a["seed"] == b["seed"]
    assert torch.equal(a["grad"], b["grad"])
    torch.manual_seed(100)                                   # rank 0's initial weights
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 1))
    g = torch.Generator().manual_seed(1)
    x = torch.randn(n, 6, generator=g)
    R = torch.randn(n, generator=g)
    loss = ((R - R.mean()) * model(x).squeeze(1)).mean()
    loss.backward()
    want = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    assert torch.allclose(a["r_mean"], R.mean(), atol=1e-7)
    assert torch.allclose(a["grad"], want, atol=1e-6)       # global-count weighting: ragged shards exact


def test_main_reads_ini_positionally():
    import configparser
    import main
    c = configparser.RawConfigParser()
    c.read(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "environment.ini"))
    assert main.section_values(c, "QWS", "ML", ["main.py", "QWS", "ML"]) == [2, 2, 128, 20, 0.0, 0.001, 10]
    assert main.section_values(c, "Normal", "ML", ["main.py", "Normal", "ML"])[1] == 4
    low = main.section_values(c, "QWS", "PNLow", ["main.py", "QWS", "PNLow", "6"])
    assert low[:9] == [0, 1, 47, 1, 5, 256, 0, 10, 1] and low[-1] == 6 and low[11] == 1e-4
    high = main.section_values(c, "Normal", "PNHigh", ["main.py", "Normal", "PNHigh", "3", "7"])
    assert high[2:5] == [50, 1, 10] and high[-2:] == [7, 3]          # argv[3] -> epochPNLow, argv[4] -> epochML
    assert main.section_values(c, "QWS", "ML+2PN", ["main.py", "QWS", "ML+2PN"]) == [47, -1]


def test_reference_import_paths_resolve():
    import src.models.modelPN as m1
    import src.models.modelML as m2
    import src.models.trainPNHigh as t
    import src.ML2PN as e
    import src.loadData as l
    assert {"Attention", "PointerNet", "CombinatorialRL", "reward"} <= set(dir(m1))
    assert {"Net", "NodeEncoder", "EdgeEncoder"} <= set(dir(m2))
    assert hasattr(t, "PNHigh") and hasattr(e, "check") and hasattr(l, "loadDataPN")


def _woa_worker(rank, world, port, out, golden):
    """ESWOA instances are independent: each rank runs its contiguous block, no collective."""
    import copy, json
    import numpy as np
    from gnnpn_sc_b200.WOA import PhiloxRng, run_many
    from oracle import woa_oracle as wo
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    probs = []
    for name in ("woa_small", "woa_tight", "woa_seeded"):
        with open(os.path.join(golden, name + ".json")) as f:
            inp = json.load(f)["input"]
        probs.append(([[tuple(s) for s in c] for c in inp["services"]], inp["constraints"], inp["solution"]))
    lo, hi = parallel.shard_range(len(probs))
    res = run_many(copy.deepcopy(probs[lo:hi]), popSize=8, MAX_Iter=12, fitness=wo.CpuFitness,
                   rngs=[PhiloxRng(50 + k) for k in range(lo, hi)])
    mine = torch.tensor([r[0] for r in res] + [0.0] * (2 - (hi - lo)), dtype=torch.float64)      # pad ragged shard
    parts = [torch.empty(2, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(parts, mine)                                                          # reporting only
    if rank == 0:
        torch.save(torch.cat(parts)[: len(probs)], out)
    dist.destroy_process_group()


def test_woa_instances_shard_without_communication(tmp_path):
    import copy, json
    from gnnpn_sc_b200.WOA import PhiloxRng, run_many
    from oracle import woa_oracle as wo
    golden = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    out = str(tmp_path / "woa.pt")
    mp.spawn(_woa_worker, args=(2, _free_port(), out, golden), nprocs=2, join=True)
    got = torch.load(out)
    probs = []
    for name in ("woa_small", "woa_tight", "woa_seeded"):
        with open(os.path.join(golden, name + ".json")) as f:
            inp = json.load(f)["input"]
        probs.append(([[tuple(s) for s in c] for c in inp["services"]], inp["constraints"], inp["solution"]))
    want = run_many(copy.deepcopy(probs), popSize=8, MAX_Iter=12, fitness=wo.CpuFitness, rngs=[PhiloxRng(50 + k) for k in range(3)])
    assert got.tolist() == [w[0] for w in want]
