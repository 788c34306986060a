#!/usr/bin/env python
"""Upper bound on what perfect lane compaction could buy the thread-per-instance BDF kernel: time the
Robertson sweep (a) as is, (b) with every warp's 32 instances identical, (c) with all instances identical.
   python tools/exp_uniform.py [batch]"""
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
base = sweeps.robertson_sweep(np.arange(B))
def morton_order(p, bits=7):
    """instances ordered along a Z-curve through (log k1, log k2, log k3): neighbours in the batch are neighbours in
    parameter space"""
    q = np.log(p)
    q = (q - q.min(axis=0)) / (q.max(axis=0) - q.min(axis=0) + 1e-300)
    q = np.minimum((q * (1 << bits)).astype(np.uint64), (1 << bits) - 1)
    code = np.zeros(len(p), dtype=np.uint64)
    for b in range(bits):
        for j in range(p.shape[1]):
            code |= ((q[:, j] >> np.uint64(b)) & np.uint64(1)) << np.uint64(b * p.shape[1] + j)
    return np.argsort(code, kind="stable")


cases = {"sweep": base, "sorted_by_k1": base[np.argsort(base[:, 0])], "morton_sorted": base[morton_order(base)], "warp_uniform": base[(np.arange(B) // 32) * 32], "block_uniform": base[(np.arange(B) // 128) * 128],
         "all_identical": np.repeat(base[:1], B, axis=0)}
for name, p in cases.items():
    prob = ds.OdeBuilder().rhs_implicit("robertson_ode").p(p).rtol(1e-4).atol([1e-8, 1e-14, 1e-6]).build()
    solver = prob.bdf()
    ms = []
    for _ in range(3):
        solver.solve_dense(sweeps.ROBERTSON_T_EVAL)
        ms.append(solver.last_integrator_ms())
    st = solver.statistics_array()
    nli = int(st[:, 8].sum())
    print(json.dumps({"case": name, "batch": B, "kernel_ms": min(ms), "nli": nli, "newton_iters_per_s": nli / min(ms) * 1e3,
                      "steps_mean": float(st[:, 6].mean()), "ns_per_newton_iter_per_sm_lane": min(ms) * 1e6 / nli * 148 * 384}))
