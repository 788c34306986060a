#!/usr/bin/env python
"""Warp-scheduler occupancy of the BDF lane kernel on the Robertson sweep.  Needs a library built with
   DSB_LIB_TAG=prof DSB_NVCC_EXTRA=-DDSB_LANE_PROFILE python -m diffsol_b200.build
   DSB_LIB_TAG=prof python tools/lane_profile.py [batch]"""
import ctypes
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
solver = ds.OdeBuilder().rhs_implicit("robertson_ode").p(p).rtol(1e-4).atol([1e-8, 1e-14, 1e-6]).build().bdf()
solver.solve_dense(sweeps.ROBERTSON_T_EVAL)
w = np.zeros(31, dtype=np.uint64)
capi.check(capi.lib().dsb_batch_debug_words(solver._b, ctypes.c_void_p(w.ctypes.data)))
w = w.astype(np.float64)
trips = w[0]
out = {"batch": B, "kernel_ms": solver.last_integrator_ms(), "warp_trips": trips, "active_lanes_per_trip": w[1] / trips,
       "lanes_waiting_for_slow_group_per_trip": w[2] / trips}
for name, k in (("slow_group", 3), ("predict_chain", 5), ("newton", 7), ("post", 9)):
    out[name] = {"runs_per_trip": w[k] / trips, "lanes_per_run": w[k + 1] / max(w[k], 1)}
print(json.dumps(out))
