#!/usr/bin/env python
"""One pass of a BASELINE configuration for ncu captures: python tools/prof_config.py spm|heat256 [batch]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import diffsol_b200 as ds  # noqa: E402
from diffsol_b200 import capi, sweeps  # noqa: E402

capi.require_device()
which = sys.argv[1]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
idx = np.arange(B)
if which == "spm":
    p = (0.6 + 0.8 * sweeps.uniform(idx, 0)).reshape(-1, 1)
    solver = ds.OdeBuilder().rhs_implicit("spm").p(p).build().bdf()
    t_eval = np.arange(1, 13) * 300.0
else:
    n = int(which[4:])
    p = np.stack([1.0 + sweeps.uniform(idx, 0), 0.1 + 0.3 * sweeps.uniform(idx, 1), 0.6 + 0.3 * sweeps.uniform(idx, 2)], axis=1)
    solver = ds.OdeBuilder().rhs_implicit("heat1d_dae_%d" % n).p(p).rtol(1e-6).atol(1e-6).build().bdf()
    t_eval = np.arange(1, 101) / 100.0 * 0.99
for _ in range(2):
    solver.solve_dense(t_eval)
    print("integrator ms", solver.last_integrator_ms())
