"""CPU oracle for the ESWOA fitness -- TEST INFRASTRUCTURE ONLY (tests/ and the generating script import it; the
product path in gnnpn_sc_b200/WOA.py never does).

Restates ``ESWOA.calc`` (src/baselines/WOA.py:87-105) in numpy float64, operation for operation, and wraps it as a
fitness-backend factory with the interface of ``gnnpn_sc_b200.WOA._GpuFitness`` so the host-side search logic can be
tested on a machine without a GPU.  Pinned by ``tests/golden/woa_*.json``: fixtures produced by executing the real
reference class (oracle/make_golden_woa.py)."""
from __future__ import annotations

import numpy as np


def calc(services, constraints):
    """(violate, objFunc) of one composition: rows of (q0, q1, q2, q3)  (WOA.py:87-105)."""
    indicator = [np.array([services[i][j] for i in range(len(services))]) for j in range(4)]
    con_values = [np.cumprod(indicator[i + 2])[-1] for i in range(2)]                     # WOA.py:93
    violate = 0
    for i in range(len(constraints)):
        for c in constraints[i]:
            if con_values[i] < c[-2] or con_values[i] > c[-1]:                            # WOA.py:96
                violate += 1
    service_num = sum(1 for s in services if s[0] > 0)                                   # WOA.py:99-101
    obj = (np.sum(indicator[0]) / service_num + 1 - np.min(indicator[1])) / 2            # WOA.py:103
    return violate, float(obj)


class CpuFitness:
    """Same call interface as the GPU backend: blocks of positions (int arrays [m, K], local index per category with
    Python index semantics) per instance -> flat (violate, objFunc, fitness) arrays."""

    def __init__(self, instances):
        self.inst = list(instances)

    def __call__(self, which, blocks):
        viol, obj, fit = [], [], []
        for w, blk in zip(which, blocks):
            it = self.inst[w]
            for pos in blk:
                rows = [it.services[c][int(v)] for c, v in enumerate(pos)]
                v, o = calc(rows, it.constraints)
                viol.append(v); obj.append(o); fit.append(v + o)                         # WOA.py:59,78,119,151
        return np.asarray(viol, dtype=np.int32), np.asarray(obj, dtype=np.float64), np.asarray(fit, dtype=np.float64)
