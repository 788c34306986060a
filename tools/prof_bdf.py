#!/usr/bin/env python
"""One BDF pass over the Robertson sweep for ncu captures: python tools/prof_bdf.py [batch] [passes]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import diffsol_b200 as ds  # noqa: E402
from diffsol_b200 import capi, sweeps  # noqa: E402

capi.require_device()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 2
p = sweeps.robertson_sweep(np.arange(B))
solver = ds.OdeBuilder().rhs_implicit("robertson_ode").p(p).rtol(1e-4).atol([1e-8, 1e-14, 1e-6]).build().bdf()
for _ in range(passes):
    solver.solve_dense(sweeps.ROBERTSON_T_EVAL)
    print("integrator ms", solver.last_integrator_ms())
