"""Synthetic QWS-shaped / Normal-shaped workloads (the reference ships no data,
`data/.gitkeep` only; layouts follow SURVEY Appendix A and 8d).

Everything is drawn from ``numpy.random.default_rng`` (PCG64) so the same seed
gives the same bytes on every box and every torch version.
"""
from __future__ import annotations

import json
import os
from typing import Dict, List, Tuple

import numpy as np
import torch

QWS = dict(K=47, N=5, S=2507)          # environment.ini [QWS-PNLow]: serCategory 47, serNumber 5
NORMAL = dict(K=50, N=10, S=2500)      # environment.ini [Normal-PNLow]: 50 x 10


def _qos(rng, shape, dist: str) -> np.ndarray:
    """4 trailing QoS attrs per service: q0 (minimise), q1 (maximise), q2/q3 multiplicative."""
    if dist == "normal":
        q01 = np.clip(rng.normal(0.5, 0.15, size=shape + (2,)), 0.01, 0.99)
        q23 = np.clip(rng.normal(0.95, 0.02, size=shape + (2,)), 0.90, 0.999)
    else:
        q01 = rng.uniform(0.0, 1.0, size=shape + (2,))
        q23 = rng.uniform(0.90, 1.0, size=shape + (2,))
    return np.concatenate([q01, q23], axis=-1).astype(np.float32)


def pn_instances(n: int, K: int, N: int, seed: int = 1234, dist: str = "qws",
                 neutral_frac: float = 0.2) -> torch.Tensor:
    """PN input rows ``fp32 [n, K*N, 8]`` = ``[q0,q1,q2,q3,g1lo,g1hi,g2lo,g2hi]``.

    Layout produced by the reference's ``loadDataPN`` + ``SCDataset``
    (loadData.py:130-148, trainPNLow.py:26-28): the four global bounds are
    non-zero only on the N rows of category 0; a category the request does not
    use is N copies of the neutral row ``[0,1,1,1,0,0,0,0]``.
    """
    rng = np.random.default_rng(seed)
    x = np.zeros((n, K, N, 8), dtype=np.float32)
    x[..., :4] = _qos(rng, (n, K, N), dist)
    neutral = rng.random((n, K)) < neutral_frac
    x[neutral, :, :4] = np.array([0, 1, 1, 1], dtype=np.float32)
    used = (~neutral).sum(axis=1).astype(np.float32)                  # [n]
    centre = np.float32(0.95) ** used
    lo = centre[:, None] * rng.uniform(0.6, 1.05, size=(n, 2)).astype(np.float32)
    bounds = np.stack([lo[:, 0], np.ones(n, np.float32), lo[:, 1], np.ones(n, np.float32)], axis=1)
    x[:, 0, :, 4:] = bounds[:, None, :].astype(np.float32)
    return torch.from_numpy(x.reshape(n, K * N, 8))


# --------------------------------------------------------------------------
# ML stage: the five JSON files of data/<dataset>/ 
# --------------------------------------------------------------------------
def category_sizes(K: int, S: int) -> List[int]:
    base, extra = divmod(S, K)
    return [base + (1 if k < extra else 0) for k in range(K)]


def ml_dataset(n_instances: int = 4000, K: int = 47, S: int = 2507, seed: int = 0,
               dist: str = "qws", min_tasks: int = 10, n_attrs: int = 9) -> Dict[str, object]:
    """Returns the python objects the reference loads with json (loadData.py:17-24,81).

    service rows carry ``n_attrs`` values of which the loaders read the last 4
    (loadData.py:40).  Request graph: node 0 = global-constraint node, then m
    task nodes; edges = bidirectional chain over the tasks + global<->task.
    Labels: one feasible service per requested task, Zipf(1.1)-ranked inside
    its category so popular services become hubs of the co-usage graph.
    """
    rng = np.random.default_rng(seed)
    sizes = category_sizes(K, S)
    offs = np.concatenate([[0], np.cumsum(sizes)])
    service_feature: Dict[str, List[List[float]]] = {}
    qos_by_cat = []
    for k in range(K):
        q = _qos(rng, (sizes[k],), dist)
        pad = rng.uniform(0, 1, size=(sizes[k], n_attrs - 4)).astype(np.float32)
        service_feature[str(k + 1)] = np.concatenate([pad, q], axis=1).astype(float).round(6).tolist()
        qos_by_cat.append(np.asarray(service_feature[str(k + 1)], dtype=np.float64)[:, -4:])

    nodefeatures, edge_indices, labels, min_cost = [], [], [], []
    for _ in range(n_instances):
        m = int(rng.integers(min_tasks, K + 1))
        tasks = np.sort(rng.choice(K, size=m, replace=False))
        nodes, chosen = [], []
        for k in tasks:
            q = qos_by_cat[k]
            # local bounds on q2 ("cost") and q3 ("quality"), wide enough to keep >=1 service feasible
            lo1, lo2 = float(np.quantile(q[:, 2], rng.uniform(0, .3))), float(np.quantile(q[:, 3], rng.uniform(0, .3)))
            feas = np.nonzero((q[:, 2] >= lo1) & (q[:, 3] >= lo2))[0]
            rank = min(int(rng.zipf(1.1)) - 1, len(feas) - 1)
            s = int(feas[rank % len(feas)])
            chosen.append((int(k), s))
            onehot = [0] * (K + 1)
            onehot[k + 1] = 1
            nodes.append(onehot + [round(float(rng.uniform()), 6), round(lo1, 6), 1.0,
                                   round(float(rng.uniform()), 6), round(lo2, 6), 1.0])
        sel = np.array([qos_by_cat[k][s] for k, s in chosen])
        p2, p3 = float(np.prod(sel[:, 2])), float(np.prod(sel[:, 3]))
        glob = [1] + [0] * K + [0.0, round(p2 * float(rng.uniform(.85, 1.0)), 6), 1.0,
                                0.0, round(p3 * float(rng.uniform(.85, 1.0)), 6), 1.0]
        nodes = [glob] + nodes
        src, dst = [], []
        for t in range(1, m + 1):
            src += [0, t]
            dst += [t, 0]
            if t < m:
                src += [t, t + 1]
                dst += [t + 1, t]
        lab = [0] * S
        for k, s in chosen:
            lab[int(offs[k]) + s] = 1
        nodefeatures.append(nodes)
        edge_indices.append([src, dst])
        labels.append(lab)
        min_cost.append(float(0.5 * (sel[:, 0].mean() + 1 - sel[:, 1].min())))
    return {"nodefeatures": nodefeatures, "edge_indices": edge_indices, "labels": labels,
            "serviceFeature": service_feature, "minCostList": min_cost}


def write_dataset(root: str, name: str, ds: Dict[str, object]) -> str:
    """Writes ``<root>/data/<name>/*.data`` exactly as loadData.py:17-24,81 reads them."""
    d = os.path.join(root, "data", name)
    os.makedirs(d, exist_ok=True)
    for key, fn in (("nodefeatures", "nodefeatures.data"), ("edge_indices", "edge_indices.data"),
                    ("labels", "labels.data"), ("serviceFeature", "serviceFeature.data"),
                    ("minCostList", "minCostList.data")):
        with open(os.path.join(d, fn), "w") as f:
            json.dump(ds[key], f)
    return d


def random_graph_csr_inputs(n_nodes: int, n_edges: int, seed: int = 7, skew: float = 0.0,
                            device="cpu") -> Tuple[torch.Tensor, torch.Tensor]:
    """Aggregation micro-benchmark graph: ``edge_index int64 [2,E]``
    with uniform destinations and uniform (``skew=0``) or Zipf-skewed sources, plus
    asymmetric positive weights."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    dst = torch.randint(0, n_nodes, (n_edges,), generator=g)
    if skew > 0:
        u = torch.rand(n_edges, generator=g, dtype=torch.float64)
        src = (n_nodes * u.pow(1.0 + skew)).long().clamp_(max=n_nodes - 1)
    else:
        src = torch.randint(0, n_nodes, (n_edges,), generator=g)
    w = torch.rand(n_edges, generator=g) + 0.05
    return torch.stack([src, dst]).to(device), w.to(device)
