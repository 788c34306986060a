#!/usr/bin/env python
"""Experiment: how much of the headline kernel's lane divergence goes away when the lanes of a warp integrate SIMILAR
instances?  The Robertson sweep's parameters are pseudo-random in the instance index; here the same 10^6 instances are handed
to the kernel in different orders (the arithmetic of an instance does not depend on its neighbours): index order, sorted by
one parameter, and along a Morton (Z-order) curve through the quantised (log k1, log k2, log k3) cube.
   python tools/exp_sorted_schedule.py [batch]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import diffsol_b200 as ds  # noqa: E402
from diffsol_b200 import capi, sweeps  # noqa: E402

capi.require_device()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
p = sweeps.robertson_sweep(np.arange(B))
tol = sweeps.ROBERTSON_ODE_TOL


def morton(p, bits=7):
    q = np.log10(p / np.array([0.04, 1.0e4, 3.0e7])) + 0.5
    q = np.clip((q * (1 << bits)).astype(np.int64), 0, (1 << bits) - 1)
    key = np.zeros(len(p), dtype=np.int64)
    for b in range(bits):
        for j in range(3):
            key |= ((q[:, j] >> b) & 1) << (3 * b + j)
    return key


orders = {"index": np.arange(B), "sorted_k1": np.argsort(p[:, 0], kind="stable"), "sorted_k3": np.argsort(p[:, 2], kind="stable"),
          "morton": np.argsort(morton(p), kind="stable")}
out = {}
for name, order in orders.items():
    s = ds.OdeBuilder().rhs_implicit("robertson_ode").p(p[order]).rtol(tol["rtol"]).atol(tol["atol"]).build().bdf()
    best = None
    for it in range(3):
        ys = s.solve_dense(sweeps.ROBERTSON_T_EVAL)
        if it > 0 and (best is None or s.last_kernel_ms() < best):
            best = s.last_kernel_ms()
    out[name] = {"kernel_ms": best, "newton_iters": int(s.statistics_array()[:, 8].sum())}
    del s
print(json.dumps(out))
