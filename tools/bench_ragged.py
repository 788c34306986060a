#!/usr/bin/env python
"""Time `solve(final_time)` (every internal step, ragged per instance) on one GPU: the counting pass, the writing pass and
the whole host call.   python tools/bench_ragged.py [batch] [bdf|tr_bdf2|esdirk34]"""
import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import diffsol_b200 as ds  # noqa: E402
from diffsol_b200 import capi, sweeps  # noqa: E402

capi.require_device()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
method = sys.argv[2] if len(sys.argv) > 2 else "bdf"
p = sweeps.robertson_sweep(np.arange(B))
tol = sweeps.ROBERTSON_ODE_TOL
solver = getattr(ds.OdeBuilder().rhs_implicit("robertson_ode").p(p).rtol(tol["rtol"]).atol(tol["atol"]).build(), method)()
best = None
for it in range(3):
    t0 = time.perf_counter()
    ys, ts, off = solver.solve(1.0e4)
    dt = (time.perf_counter() - t0) * 1e3
    if it > 0 and (best is None or dt < best):
        best = dt
L = capi.lib()
total = ctypes.c_int64()
solver.set_params()
t0 = time.perf_counter()
capi.check(L.dsb_batch_solve_count(solver._b, solver.method, 1.0e4, ctypes.byref(total)))
count_ms = (time.perf_counter() - t0) * 1e3
print(json.dumps({"method": method, "batch": B, "columns_total": int(total.value), "columns_per_instance_mean": total.value / B,
                  "host_call_ms": best, "counting_pass_ms": count_ms, "result_MB": (ys.nbytes + ts.nbytes) / 1e6}))
