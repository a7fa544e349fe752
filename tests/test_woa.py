"""ESWOA fine-tuning stage: host search logic and GPU fitness against fixtures produced by executing the
real reference class (oracle/make_golden_woa.py -> tests/golden/woa_*.json).  Everything is float64 and must be
bit-identical: fitness values, the best-so-far trajectory, the final position."""
import copy
import json
import os

import numpy as np
import pytest

from oracle import woa_oracle as wo

CASES = ["woa_small", "woa_seeded", "woa_qws_shaped", "woa_tight"]


def _load(golden_dir, name):
    with open(os.path.join(golden_dir, name + ".json")) as f:
        d = json.load(f)
    inp = d["input"]
    services = [[tuple(s) for s in cat] for cat in inp["services"]]
    return inp, services, d["reference"]


def _run(inp, services, fitness):
    from gnnpn_sc_b200.WOA import ESWOA
    np.random.seed(inp["seed"])
    m = ESWOA(copy.deepcopy(services), copy.deepcopy(inp["constraints"]), copy.deepcopy(inp["solution"]),
              popSize=inp["popSize"], MAX_Iter=inp["MAX_Iter"], fitness=fitness)
    after_init = (m.bestFitness, list(m.bestPops))
    best, sol = m.start()
    return m, after_init, best, sol


def _check(m, after_init, best, sol, ref):
    assert m.initFitness == ref["initFitness"]
    assert after_init[0] == ref["bestFitness_after_init"] and after_init[1] == ref["bestPops_after_init"]
    assert m.bestFitnesses == ref["bestFitnesses"]                     # the whole trajectory, bit for bit
    assert best == ref["bestFitness"] and [int(x) for x in m.bestPops] == ref["bestPops"]
    assert [list(map(float, r)) for r in sol] == ref["bestSolutions"]
    assert [[list(map(float, s)) for s in c] for c in m.services] == ref["services_after_init"]


@pytest.mark.parametrize("name", CASES)
def test_oracle_calc_matches_reference_probes(name, golden_dir):
    inp, services, ref = _load(golden_dir, name)
    svc = [[tuple(s) for s in c] for c in ref["services_after_init"]]
    for pr in ref["probes"]:
        rows = [svc[c][v] for c, v in enumerate(pr["pos"])]
        v, o = wo.calc(rows, inp["constraints"])
        assert v == pr["violate"] and o == pr["objFunc"]


@pytest.mark.parametrize("name", CASES)
def test_host_search_replays_reference_run(name, golden_dir):
    """Speculative batched evaluation + in-order best-so-far replay == the reference's one-whale-at-a-time loop
    (same np.random stream, same list aliasing), with the CPU oracle as the fitness backend."""
    inp, services, ref = _load(golden_dir, name)
    _check(*_run(inp, services, wo.CpuFitness), ref)


def test_lockstep_batch_equals_single_runs(golden_dir):
    from gnnpn_sc_b200.WOA import ESWOA, run_many
    probs, singles = [], []
    for k, name in enumerate(CASES):
        inp, services, _ = _load(golden_dir, name)
        probs.append((copy.deepcopy(services), copy.deepcopy(inp["constraints"]), copy.deepcopy(inp["solution"])))
        m = ESWOA(copy.deepcopy(services), copy.deepcopy(inp["constraints"]), copy.deepcopy(inp["solution"]), popSize=10,
                  MAX_Iter=25, rng=np.random.RandomState(100 + k), fitness=wo.CpuFitness)
        m.start()
        singles.append((m.bestFitness, m.bestFitnesses))
    many = run_many(probs, popSize=10, MAX_Iter=25, seeds=[100 + k for k in range(len(CASES))], fitness=wo.CpuFitness)
    for (bf, traj), (mbf, _, mtraj) in zip(singles, many):
        assert bf == mbf and traj == mtraj


def test_multi_constraint_lists_are_rejected_loudly():
    from gnnpn_sc_b200.WOA import ESWOA
    with pytest.raises(NotImplementedError):
        ESWOA([[(0.5, 0.5, 0.95, 0.95)]], [[[0.1, 1.0], [0.2, 1.0]], [[0.1, 1.0]]], popSize=2, MAX_Iter=1, fitness=wo.CpuFitness)


def test_file_level_driver_contract(tmp_path):
    """WOA(...).start(): allActions{epoch}.txt + data/<ds>/*.data in, solutions/WOA/<ds>/ML+2PN+WOA.txt out."""
    from gnnpn_sc_b200 import WOA as W
    g = np.random.default_rng(3)
    K, n = 4, 8
    svc = {str(c + 1): [[0.0] * 5 + [float(g.uniform(0.05, 1)), float(g.uniform(0.05, 1)), float(g.uniform(0.9, 1)),
                                     float(g.uniform(0.9, 1))] for _ in range(6)] for c in range(K)}
    nodef, mincost = [], []
    for _ in range(n):
        nodes = [[1] + [0] * K + [0, 0.5, 1.0, 0, 0.5, 1.0]]                       # global node: bounds on the products
        for c in range(1, K + 1):
            nodes.append([0] * c + [1] + [0] * (K - c) + [0, 0.0, 1.0, 0, 0.0, 1.0])
        nodef.append(nodes); mincost.append(0.4)
    root = str(tmp_path)
    os.makedirs(os.path.join(root, "data", "toy")); os.makedirs(os.path.join(root, "solutions", "PNHigh", "toy"))
    for name, obj in (("nodefeatures.data", nodef), ("serviceFeature.data", svc), ("minCostList.data", mincost),
                      ("labels.data", [[0] * (6 * K)] * n)):
        with open(os.path.join(root, "data", "toy", name), "w") as f:
            json.dump(obj, f)
    n_test = n - n // 4 * 3
    actions = [[svc[str(c + 1)][int(g.integers(0, 6))][-4:] + [0, 0, 0, 0] for _ in range(n_test)] for c in range(K)]
    actions[2][0] = [0, 1, 1, 1, 0, 0, 0, 0]                                        # a neutral row (dropped)
    with open(os.path.join(root, "solutions", "PNHigh", "toy", "allActions3.txt"), "w") as f:
        json.dump(actions, f)
    np.random.seed(0)
    out = W.WOA("toy", K, 0, 1, 0, 0, 5, 0, 3, 6, 5, root=root, fitness=wo.CpuFitness).start()
    with open(os.path.join(root, "solutions", "WOA", "toy", "ML+2PN+WOA.txt")) as f:
        saved = json.load(f)
    assert saved["quality"] == out["quality"] and len(out["quality"]) == n_test
    assert all(q > 0 for q in out["quality"]) and set(saved) == {"quality", "time", "averageQ", "averageT"}


# ----------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_gpu_fitness_bitwise_vs_oracle():
    import torch
    from gnnpn_sc_b200 import ops
    g = np.random.default_rng(5)
    for K, P in ((3, 50), (47, 400), (100, 300), (200, 64)):              # K = 200 takes numpy's recursive pairwise path
        sizes = g.integers(2, 40, K)
        base = np.concatenate([[0], np.cumsum(sizes)])
        q = np.stack([g.uniform(-0.2, 1, base[-1]), g.uniform(0, 1, base[-1]), g.uniform(0.9, 1, base[-1]),
                      g.uniform(0.9, 1, base[-1])], 1)
        klen = g.integers(1, K + 1, P).astype(np.int32)
        local = np.stack([g.integers(0, sizes) for _ in range(P)])
        idx = (base[:-1][None, :] + local).astype(np.int32)
        bounds = np.stack([g.uniform(0, 0.5, P), g.uniform(0.5, 1, P), g.uniform(0, 0.5, P), g.uniform(0.5, 1, P)], 1)
        viol, obj, fit = ops.woa_fitness(torch.from_numpy(q).cuda(), torch.from_numpy(idx).cuda(),
                                         torch.from_numpy(bounds).cuda(), torch.from_numpy(klen).cuda())
        viol, obj, fit = viol.cpu().numpy(), obj.cpu().numpy(), fit.cpu().numpy()
        for p in range(P):
            rows = [tuple(q[idx[p, k]]) for k in range(klen[p])]
            cons = [[[bounds[p, 0], bounds[p, 1]]], [[bounds[p, 2], bounds[p, 3]]]]
            v, o = wo.calc(rows, cons)
            assert viol[p] == v and (obj[p] == o or (np.isnan(obj[p]) and np.isnan(o))), (K, p, obj[p], o)
            assert fit[p] == v + o or np.isnan(fit[p])


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gpu_eswoa_replays_reference_run(name, golden_dir):
    inp, services, ref = _load(golden_dir, name)
    from gnnpn_sc_b200 import _lib
    before = _lib.launch_count()
    _check(*_run(inp, services, None), ref)                                  # default backend = the CUDA kernel
    assert _lib.launch_count() > before


@pytest.mark.gpu
def test_gpu_lockstep_many_instances(golden_dir):
    from gnnpn_sc_b200.WOA import run_many
    probs = []
    for name in CASES * 8:
        inp, services, _ = _load(golden_dir, name)
        probs.append((copy.deepcopy(services), copy.deepcopy(inp["constraints"]), copy.deepcopy(inp["solution"])))
    seeds = list(range(len(probs)))
    gpu = run_many(copy.deepcopy(probs), popSize=12, MAX_Iter=20, seeds=seeds)
    cpu = run_many(copy.deepcopy(probs), popSize=12, MAX_Iter=20, seeds=seeds, fitness=wo.CpuFitness)
    for a, b in zip(gpu, cpu):
        assert a[0] == b[0] and a[2] == b[2]


def test_load_data_other_matches_reference(tmp_path, golden_dir):
    """loadDataOther + the addS candidate filter (plain, dominance-reduced, reduced with protected rows) against what
    the reference's own functions returned on the same toy dataset (oracle/make_golden_woa.py: loader_fixture)."""
    from gnnpn_sc_b200.WOA import loadDataOther
    with open(os.path.join(golden_dir, "woa_loader.json")) as f:
        d = json.load(f)
    root = str(tmp_path)
    os.makedirs(os.path.join(root, "data", "toy"))
    for name, obj in d["data"].items():
        with open(os.path.join(root, "data", "toy", name), "w") as f:
            json.dump(obj, f)
    sset = [{tuple(r) for r in rows} for rows in d["sset"]]
    for key, (reduct, ss) in {"plain": (False, None), "reduct": (0.55, None), "reduct_protected": (0.55, sset)}.items():
        feats, cons, mc = loadDataOther("toy", reduct, sSetList=ss, train=False, root=root)
        ref = d["reference"][key]
        assert [[[list(s) for s in c] for c in inst] for inst in feats] == ref["features"], key
        assert cons == ref["constraints"] and mc == ref["minCost"]


def test_philox_host_engine_is_deterministic_and_context_addressed(golden_dir):
    """The counter-based generator gives the same run regardless of how instances are grouped (draws are addressed by
    (seed, iteration, phase, whale, slot), not by call order across instances)."""
    from gnnpn_sc_b200.WOA import PhiloxRng, run_many
    probs = []
    for name in CASES:
        inp, services, _ = _load(golden_dir, name)
        probs.append((services, inp["constraints"], inp["solution"]))
    mk = lambda: copy.deepcopy(probs)
    a = run_many(mk(), popSize=10, MAX_Iter=15, fitness=wo.CpuFitness, rngs=[PhiloxRng(7 + k) for k in range(4)])
    b = [run_many([mk()[k]], popSize=10, MAX_Iter=15, fitness=wo.CpuFitness, rngs=[PhiloxRng(7 + k)])[0] for k in range(4)]
    assert [x[0] for x in a] == [x[0] for x in b] and [x[2] for x in a] == [x[2] for x in b]
    u = PhiloxRng(1); u.at(3, 2, 5)
    v = PhiloxRng(1); v.at(3, 2, 5)
    assert u.random() == v.random() and 0.0 <= u.random() < 1.0


@pytest.mark.gpu
def test_device_search_equals_host_engine_bitwise(golden_dir):
    """gnnpn_woa_search_f64 (whole search in one launch) == the host engine driven by the same Philox draws: best-so-far
    trajectory, final fitness and final position, bit for bit; the host engine itself replays the reference
    (test_gpu_eswoa_replays_reference_run)."""
    from gnnpn_sc_b200.WOA import PhiloxRng, run_many, run_many_device
    probs = []
    for name in CASES * 3:
        inp, services, _ = _load(golden_dir, name)
        probs.append((services, inp["constraints"], inp["solution"]))
    seeds = [1000 + 17 * k for k in range(len(probs))]
    for pop, iters in ((12, 40), (50, 30), (128, 12)):
        dev = run_many_device(copy.deepcopy(probs), popSize=pop, MAX_Iter=iters, seeds=seeds)
        host = run_many(copy.deepcopy(probs), popSize=pop, MAX_Iter=iters, rngs=[PhiloxRng(s) for s in seeds])
        for k, (d, h) in enumerate(zip(dev, host)):
            assert d[2] == h[2], (pop, iters, k, next(i for i, (x, y) in enumerate(zip(d[2], h[2])) if x != y))
            assert d[0] == h[0] and [tuple(r) for r in d[1]] == [tuple(r) for r in h[1]]


@pytest.mark.gpu
def test_file_level_driver_device_engine(tmp_path):
    """WOA(..., engine="device"): same files in and out, every instance searched in one launch."""
    from gnnpn_sc_b200 import WOA as W
    g = np.random.default_rng(3)
    K, n = 4, 8
    svc = {str(c + 1): [[0.0] * 5 + [float(g.uniform(0.05, 1)), float(g.uniform(0.05, 1)), float(g.uniform(0.9, 1)),
                                     float(g.uniform(0.9, 1))] for _ in range(6)] for c in range(K)}
    nodef = []
    for _ in range(n):
        nodes = [[1] + [0] * K + [0, 0.5, 1.0, 0, 0.5, 1.0]]
        for c in range(1, K + 1):
            nodes.append([0] * c + [1] + [0] * (K - c) + [0, 0.0, 1.0, 0, 0.0, 1.0])
        nodef.append(nodes)
    root = str(tmp_path)
    os.makedirs(os.path.join(root, "data", "toy"))
    for name, obj in (("nodefeatures.data", nodef), ("serviceFeature.data", svc), ("minCostList.data", [0.4] * n),
                      ("labels.data", [[0] * (6 * K)] * n)):
        with open(os.path.join(root, "data", "toy", name), "w") as f:
            json.dump(obj, f)
    out = W.WOA("toy", K, 0, 0, 0, 1, 5, 0, -1, 20, 16, root=root, engine="device").start()
    with open(os.path.join(root, "solutions", "WOA", "toy", "ESWOA.txt")) as f:
        saved = json.load(f)
    assert saved["quality"] == out["quality"] and len(out["quality"]) == n - n // 4 * 3 and all(q > 0 for q in out["quality"])
    # the search can only improve on the best member of the initial population
    from gnnpn_sc_b200.WOA import loadDataOther, run_many_device
    feats, cons, mc = loadDataOther("toy", 0, None, False, root)
    res = run_many_device([(f, c, None) for f, c in zip(feats, cons)], popSize=16, MAX_Iter=20, seeds=list(range(6, 8)))
    assert all(r[2][-1] <= r[2][0] for r in res) and [mc[6 + k] / r[0] for k, r in enumerate(res)] == out["quality"]
