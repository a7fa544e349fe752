"""Generate tests/golden/woa_*.json by EXECUTING the real reference class (src/baselines/WOA.py: ESWOA) in the build
container (/root/reference is importable here; it does not exist on the GPU box, hence committed fixtures).

    python oracle/make_golden_woa.py

Each fixture holds the synthetic inputs (candidate services per requested task, the two global constraints, an optional
seed solution), the numpy seed, and what the reference produced: initial / final best fitness, the best-so-far
trajectory ``bestFitnesses``, the final ``bestPops``, and ``calc`` of a few random compositions."""
import copy, json, os, sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
CASES = {
    # name: (tasks, candidates per task (min, max), popSize, MAX_Iter, with seed solution, numpy seed, round inputs)
    "woa_small": (6, (3, 9), 12, 40, False, 11),
    "woa_seeded": (15, (5, 30), 20, 60, True, 12),
    "woa_qws_shaped": (47, (20, 54), 30, 50, True, 13),
    "woa_tight": (9, (2, 4), 10, 80, False, 14),
}


def make_problem(K, span, seed, with_solution):
    g = np.random.default_rng(seed)
    services = []
    for _ in range(K):
        n = int(g.integers(span[0], span[1] + 1))
        q = np.stack([g.uniform(0.01, 1, n), g.uniform(0.01, 1, n), g.uniform(0.90, 1, n), g.uniform(0.90, 1, n)], 1)
        services.append([tuple(float(v) for v in row) for row in q])
    p2 = float(np.prod([np.mean([s[2] for s in cat]) for cat in services]))
    p3 = float(np.prod([np.mean([s[3] for s in cat]) for cat in services]))
    constraints = [[[p2 * 0.98, 1.0]], [[p3 * 1.01, 1.0]]]          # the second one is violated by average picks
    solution = None
    if with_solution:
        solution = [list(cat[int(g.integers(0, len(cat)))]) for cat in services]
        solution[1] = [0.123456789, 0.5, 0.95, 0.97]                 # a row that is not in its category (gets appended)
    return services, constraints, solution


def main():
    sys.path.insert(0, "/root/reference")
    from src.baselines.WOA import ESWOA                               # the real thing
    os.makedirs(OUT, exist_ok=True)
    for name, (K, span, pop, iters, with_sol, seed) in CASES.items():
        services, constraints, solution = make_problem(K, span, seed, with_sol)
        inp = {"services": copy.deepcopy(services), "constraints": copy.deepcopy(constraints),
               "solution": copy.deepcopy(solution), "popSize": pop, "MAX_Iter": iters, "seed": seed}
        np.random.seed(seed)
        m = ESWOA(copy.deepcopy(services), copy.deepcopy(constraints), copy.deepcopy(solution), popSize=pop, MAX_Iter=iters)
        init_fit, init_best = m.bestFitness, list(m.bestPops)
        best, sol = m.start()
        g = np.random.default_rng(seed + 100)
        probes = []
        for _ in range(16):
            pos = [int(g.integers(0, len(c))) for c in m.services]
            rows = [m.services[c][v] for c, v in enumerate(pos)]
            v, o, _ = m.calc(rows)
            probes.append({"pos": pos, "violate": int(v), "objFunc": float(o)})
        ref = {"initFitness": float(m.initFitness), "bestFitness_after_init": float(init_fit), "bestPops_after_init": [int(x) for x in init_best],
               "bestFitness": float(best), "bestFitnesses": [float(x) for x in m.bestFitnesses],
               "bestPops": [int(x) for x in m.bestPops], "bestSolutions": [list(map(float, r)) for r in sol],
               "services_after_init": [[list(map(float, s)) for s in c] for c in m.services], "probes": probes}
        with open(os.path.join(OUT, name + ".json"), "w") as f:
            json.dump({"input": inp, "reference": ref}, f)
        print(name, "best", best, "improvements", len(set(m.bestFitnesses)))


def loader_fixture():
    """loadDataOther / addS of the real reference (src/loadData.py:155-288) on a toy dataset, without and with the
    dominance reduction (reduct = 0.55) and protected rows (sSetList)."""
    import tempfile
    sys.path.insert(0, "/root/reference")
    from src.loadData import loadDataOther
    g = np.random.default_rng(21)
    K, n, per = 5, 12, 14
    svc = {str(c + 1): [[0.0] * 5 + [float(g.uniform(0.05, 1)), float(g.uniform(0.05, 1)), float(g.uniform(0.9, 1)),
                                     float(g.uniform(0.9, 1))] for _ in range(per)] for c in range(K)}
    nodef = []
    for _ in range(n):
        nodes = [[1] + [0] * K + [0, 0.5, 1.0, 0, 0.5, 1.0]]
        for c in sorted(g.choice(np.arange(1, K + 1), size=int(g.integers(2, K + 1)), replace=False).tolist()):
            lo2, lo3 = float(g.uniform(0.9, 0.94)), float(g.uniform(0.9, 0.94))
            nodes.append([0] * c + [1] + [0] * (K - c) + [0, lo2, 1.0, 0, lo3, 1.0])
        nodef.append(nodes)
    data = {"nodefeatures.data": nodef, "serviceFeature.data": svc, "minCostList.data": [0.5] * n,
            "labels.data": [[0] * (per * K)] * n}
    n_test = n - n // 4 * 3
    sset = [set() for _ in range(n_test)]
    for i in range(n_test):
        for c in range(K):
            f = svc[str(c + 1)][int(g.integers(0, per))]
            sset[i].add(tuple(round(v, 5) for v in f[-4:]))
    out = {}
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.makedirs(os.path.join(tmp, "data", "toy"))
        for name, obj in data.items():
            with open(os.path.join(tmp, "data", "toy", name), "w") as f:
                json.dump(obj, f)
        os.chdir(tmp)
        try:
            for key, (reduct, ss) in {"plain": (False, None), "reduct": (0.55, None), "reduct_protected": (0.55, sset)}.items():
                feats, cons, mc = loadDataOther("toy", reduct, sSetList=ss, train=False)
                out[key] = {"features": feats, "constraints": cons, "minCost": mc}
        finally:
            os.chdir(cwd)
    with open(os.path.join(OUT, "woa_loader.json"), "w") as f:
        json.dump({"data": data, "sset": [sorted(map(list, x)) for x in sset], "reference": out}, f)
    print("woa_loader", {k: [len(c) for c in v["features"][0]] for k, v in out.items()})


if __name__ == "__main__":
    main()
    loader_fixture()
