"""oracle/ml2pn_oracle.py against the reference's own ``ML2PN.calc`` executed live (build container only)."""
import importlib.util
import os
import sys
import types

import numpy as np
import pytest

from oracle import ml2pn_oracle as mo


def _cases(rng, n=40, K=7):
    acts = rng.uniform(0.05, 1.0, size=(K, n, 8))
    acts[:, :, 2:4] = rng.uniform(0.9, 1.0, size=(K, n, 2))
    acts[rng.integers(0, K, 5), rng.integers(0, n, 5), 0] = 0.0                  # picks with q0 == 0 stay in the mean
    neutral = rng.random((K, n)) < 0.2
    neutral[0] = False
    acts[neutral] = np.array([0, 1, 1, 1, 0, 0, 0, 0.0])
    cons = np.stack([rng.uniform(0.5, 0.8, n), rng.uniform(0.8, 1.0, n), rng.uniform(0.5, 0.8, n),
                     rng.uniform(0.8, 1.0, n)], axis=1)
    return acts.tolist(), cons.tolist()


def test_oracle_equals_live_reference_calc(reference_available):
    if not reference_available:
        pytest.skip("reference tree not mounted")
    # src/ML2PN.py imports src.loadData at module level; only `calc` is exercised, so a stub package satisfies it
    stub = types.ModuleType("src.loadData")
    stub.loadDataPN = None
    saved = {k: sys.modules.get(k) for k in ("src", "src.loadData")}
    sys.modules["src"] = types.ModuleType("src")
    sys.modules["src.loadData"] = stub
    try:
        spec = importlib.util.spec_from_file_location("_reference_ML2PN", "/root/reference/src/ML2PN.py")
        ref = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    acts, cons = _cases(np.random.default_rng(5))
    K, n = len(acts), len(acts[0])
    got = mo.scores(acts, K, cons)
    for j in range(n):
        sol = [acts[i][j][0:4] for i in range(K) if sum(acts[i][j][0:4]) != 3]
        qos = [[s[i] for s in sol] for i in range(4)]
        want = ref.calc(qos, [cons[j][:2], cons[j][2:]])
        assert got[j] == want


def test_mean_runs_over_all_real_picks():
    acts = [[[0.0, 0.5, 1.0, 1.0, 0, 0, 0, 0]], [[0.6, 0.25, 1.0, 1.0, 0, 0, 0, 0]]]
    s = mo.scores(acts, 2, [[0.0, 2.0, 0.0, 2.0]])
    assert s[0] == 0.5 * (0.3 + 1 - 0.25)        # ESWOA.calc would divide by the one pick with q0 > 0
