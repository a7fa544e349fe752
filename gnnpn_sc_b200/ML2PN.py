"""Drop-in for ``src/ML2PN.py``: score the saved PNHigh picks against the optimum.

``check(dataset, serCategory, epoch)`` reads the same files (``allActions{epoch}.txt`` or the pretrained
stand-in, ``minCostList.data``) and prints ``epoch, mean(minCost / obj)``; the objective of every test
instance is evaluated in one launch of ``gnnpn_pn_reward_f32`` (ML2PN.py:6-12 per instance in numpy).
Neutral rows (categories the request does not use: ``[0,1,1,1]``, ML2PN.py:42) are dropped by re-packing
each instance's real picks to the front before the kernel call.
"""
from __future__ import annotations

import json
import os

import numpy as np
import torch

from . import ops
from .loadData import loadDataPN


def composition_scores(allActions, serCategory: int, constraints: np.ndarray) -> np.ndarray:
    """obj per instance (ML2PN.calc): 0.5*(mean q0 + 1 - min q1) + #violated global constraints.
    ``allActions`` = K lists of [n_test][F] chosen rows; ``constraints`` [n_test, 4] = (lo1, hi1, lo2, hi2)."""
    acts = np.asarray([a for a in allActions[:serCategory]], dtype=np.float32)      # [K, n, F]
    K, n, F = acts.shape
    rows = np.transpose(acts, (1, 0, 2)).copy()                                      # [n, K, F]
    real = rows[:, :, :4].sum(axis=2) != 3                                           # ML2PN.py:42
    out = np.zeros(n, dtype=np.float64)
    # group instances by their number of real picks so every kernel call has a fixed K
    counts = real.sum(axis=1)
    dev = torch.device("cuda")
    for k in np.unique(counts):
        if k == 0:
            continue
        sel = np.nonzero(counts == k)[0]
        packed = np.zeros((len(sel), int(k), 8), dtype=np.float32)
        for r, i in enumerate(sel):
            packed[r, :, :4] = rows[i, real[i], :4]
            packed[r, 0, 4:] = constraints[i]
        x = torch.from_numpy(packed).to(dev)
        idx = torch.arange(int(k), device=dev, dtype=torch.int32).view(-1, 1).expand(int(k), len(sel)).contiguous()
        viol, obj, _ = ops.pn_reward(x, idx)
        out[sel] = (obj.double() + viol.double()).cpu().numpy()
    return out


def check(dataset, serCategory, epoch, root="."):
    feats, _ = loadDataPN(epoch=-1, dataset=dataset, serviceNumber=1, root=root, rng=False)
    split = len(feats) // 4 * 3
    with open(os.path.join(root, "data", dataset, "minCostList.data")) as f:
        minCost = json.load(f)
    url = (os.path.join(root, "solutions", "pretrained", f"{dataset}-PNHigh.txt") if epoch == -1 else
           os.path.join(root, "solutions", "PNHigh", dataset, f"allActions{epoch}.txt"))
    with open(url) as f:
        allActions = json.load(f)
    cons = np.asarray([inst[0][5:9] for inst in feats[split:]], dtype=np.float32)
    obj = composition_scores(allActions, serCategory, cons)
    ratio = float(np.mean(np.asarray(minCost[split:split + len(obj)]) / obj))
    print(epoch, ratio)
    return ratio
