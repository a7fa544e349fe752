"""Data layer with the reference's entry points (``src/loadData.py``): ``loadData`` for the ML stage and
``loadDataPN`` for the pointer networks.  Same files, same return values and
ordering; the O(n·K²) / O(S²) Python loops of the reference are replaced by numpy array code.

This is host-side glue on either side of the hot path, not part of it.
"""
from __future__ import annotations

import json
import os
from typing import Dict, List, Optional, Sequence

import numpy as np

TRAIN_INSTANCES = 3000        # the reference hard-wires the training split (loadData.py:44,67)


def compute_inv_propesity(labels, A, B):
    """loadData.py:6-11 (result unused by TrainML, kept for signature parity)."""
    lab = np.asarray(labels)
    n = lab.shape[0]
    freqs = np.ravel(lab.sum(axis=0))
    C = (np.log(n) - 1) * np.power(B + 1, A)
    return np.ravel(1.0 + C * np.power(freqs + B, -A))


def _read(dataset: str, name: str, root: str = "."):
    with open(os.path.join(root, "data", dataset, name), "r") as f:
        return json.load(f)


def compact_node_features(nodefeatures):
    """one-hot(K+1) ⊕ 6 floats  ->  [type index] ⊕ 6 floats   (loadData.py:26-33)."""
    out = []
    for inst in nodefeatures:
        a = np.asarray(inst, dtype=np.float64)
        kind = np.argmax(a[:, :-6] == 1, axis=1)
        out.append(np.concatenate([kind[:, None].astype(np.float64), a[:, -6:]], axis=1).tolist())
    return out


def service_feature_list(serviceFeature: Dict[str, list]) -> List[List[float]]:
    """rows ``[category - first_category] ⊕ last 4 QoS`` in category-key order (loadData.py:35-40)."""
    keys = sorted(int(k) for k in serviceFeature.keys())
    rows = []
    for key in keys:
        for feat in serviceFeature[str(key)]:
            rows.append([key - keys[0]] + list(feat[-4:]))
    return rows


def cousage_graph(labels, n_train: int = TRAIN_INSTANCES):
    """Service co-usage graph of the training split (loadData.py:42-65): ``adj = LᵀL`` off the diagonal, one
    directed edge pair per co-used pair (i<j): ``i->j`` weighted ``adj/use[i]`` then ``j->i`` weighted ``adj/use[j]``."""
    L = np.asarray(labels[:n_train], dtype=np.int64)
    use = L.sum(axis=0)
    adj = L.T @ L
    np.fill_diagonal(adj, 0)
    i, j = np.nonzero(np.triu(adj, 1))
    src = np.stack([i, j], axis=1).ravel()
    dst = np.stack([j, i], axis=1).ravel()
    w = np.stack([adj[i, j] / use[i], adj[j, i] / use[j]], axis=1).ravel()
    return [src.tolist(), dst.tolist()], w.tolist()


def ml_arrays(ds: Dict[str, object]):
    """Everything ``loadData`` returns, from the already-parsed JSON objects."""
    nodefeatures = compact_node_features(ds["nodefeatures"])
    services = service_feature_list(ds["serviceFeature"])
    edge_index_service, edge_attr_service = cousage_graph(ds["labels"])
    inv = compute_inv_propesity(ds["labels"][:TRAIN_INSTANCES], 0.55, 1.5)
    return nodefeatures, services, ds["edge_indices"], edge_index_service, edge_attr_service, ds["labels"], inv


def loadData(dataset: str = "", root: str = "."):
    """loadData.py:14-69."""
    ds = {"nodefeatures": _read(dataset, "nodefeatures.data", root),
          "edge_indices": _read(dataset, "edge_indices.data", root),
          "labels": _read(dataset, "labels.data", root),
          "serviceFeature": _read(dataset, "serviceFeature.data", root)}
    return ml_arrays(ds)


# --------------------------------------------------------------------------- PN inputs from the ML ranking
def pn_rows_from_ranking(nodefeature, ranking: Sequence[int], serviceFeature: Dict[str, list],
                         ser2cat: np.ndarray, ser2pos: np.ndarray, serviceNumber: int,
                         rng: Optional[np.random.Generator] = None) -> List[List[float]]:
    """One instance of loadDataPN (loadData.py:99-150): the first ``serviceNumber`` services of every requested
    category, in ranking order, that satisfy the task's local bounds on q2/q3; padded by self-duplication;
    category 0 rows carry the four global bounds; categories the request does not use become neutral rows
    ``[cat,0,1,1,1, ...]``.  The reference shuffles each candidate list with the unseeded global numpy RNG
    (loadData.py:135); pass ``rng`` for a reproducible shuffle, or ``rng=False`` to keep ranking order."""
    K = len(serviceFeature)
    cons = np.zeros((K + 1, 8))
    used = set()
    for node in nodefeature:
        if node[0] == 1:
            cons[1:, 4:] = node[-5:-3] + node[-2:]
        else:
            idx = node[:-6].index(1)
            cons[idx, :4] = node[-5:-3] + node[-2:]
            used.add(idx)
    picked: List[List[int]] = [[] for _ in range(K)]
    seen = [set() for _ in range(K)]
    for s in ranking:
        c = int(ser2cat[s])
        if len(seen[c]) < serviceNumber:
            feat = serviceFeature[str(c + 1)][int(ser2pos[s])]
            if cons[c + 1, 0] <= feat[-2] <= cons[c + 1, 1] and cons[c + 1, 2] <= feat[-1] <= cons[c + 1, 3]:
                if s not in seen[c]:
                    seen[c].add(s)
                    picked[c].append(s)
    rows: List[List[float]] = []
    for c in range(K):
        tail = cons[c + 1, 4:].tolist() if c == 0 else [0, 0, 0, 0]
        cand = list(seen[c]) if rng is None else picked[c]         # reference: set iteration order, then shuffle
        if rng is None:
            np.random.shuffle(cand)
        elif rng is not False:
            rng.shuffle(cand)
        if (c + 1) in used and cand:          # (the reference loops forever on an empty feasible set)
            while len(cand) < serviceNumber:
                cand += cand
            for v in cand[:serviceNumber]:
                f = serviceFeature[str(c + 1)][int(ser2pos[v])]
                rows.append([c] + [f[k] for k in (-4, -3, -2, -1)] + tail)
        else:
            rows += [[c, 0, 1, 1, 1] + tail for _ in range(serviceNumber)]
    return rows


def loadDataPN(epoch: int = 7, dataset: str = "", serviceNumber: int = 5, root: str = ".", rng=None):
    """loadData.py:72-152: (PN input rows per instance, minCost per instance)."""
    nodefeatures = _read(dataset, "nodefeatures.data", root)
    serviceFeature = _read(dataset, "serviceFeature.data", root)
    minCostList = _read(dataset, "minCostList.data", root)
    if epoch >= 0:
        path = os.path.join(root, "solutions", "ML", dataset, f"testServices-epoch{epoch}.txt")
    else:
        path = os.path.join(root, "solutions", "pretrained", f"{dataset}-ML.txt")
    with open(path, "r") as f:
        testServices = json.load(f)
    ser2cat, ser2pos = [], []
    for key in serviceFeature.keys():
        ser2cat += [int(key) - 1] * len(serviceFeature[key])
        ser2pos += list(range(len(serviceFeature[key])))
    ser2cat, ser2pos = np.asarray(ser2cat), np.asarray(ser2pos)
    feats, labels = [], []
    for nodefeature, ranking, minCost in zip(nodefeatures, testServices, minCostList):
        feats.append(pn_rows_from_ranking(nodefeature, ranking, serviceFeature, ser2cat, ser2pos, serviceNumber, rng))
        labels.append(minCost)
    return feats, labels
