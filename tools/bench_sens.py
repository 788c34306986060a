#!/usr/bin/env python
"""Time the forward-sensitivity kernels on one GPU:  python tools/bench_sens.py [batch] [robertson_ode|robertson_dae|exp_decay] [bdf|tr_bdf2|esdirk34]
The headline sweep (robertson_ode, Bdf, t in [0, 1e4]) with d y / d (k1, k2, k3) beside the state and in the error test
(sens_rtol = rtol, sens_atol 1e-6).  Prints kernel ms, instances/s and Newton-it/s (state and sensitivity solves together)."""
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
model = sys.argv[2] if len(sys.argv) > 2 else "robertson_ode"
method = sys.argv[3] if len(sys.argv) > 3 else "bdf"
if model.startswith("robertson"):
    p = sweeps.robertson_sweep(np.arange(B))
    tol = sweeps.ROBERTSON_ODE_TOL if model == "robertson_ode" else dict(rtol=1e-4, atol=[1e-8, 1e-6, 1e-6])
    t_eval, nsa = sweeps.ROBERTSON_T_EVAL, 3
else:
    i = np.arange(B)
    p = np.stack([0.02 * 100.0 ** sweeps.uniform(i, 0), 0.5 + 1.5 * sweeps.uniform(i, 1)], axis=1)
    tol, t_eval, nsa = dict(rtol=1e-6, atol=[1e-6]), np.linspace(0.5, 10.0, 20), 2
builder = ds.OdeBuilder().rhs_implicit(model).p(p).rtol(tol["rtol"]).atol(tol["atol"])
# TR-BDF2's unfiltered sensitivity error estimate needs millions of steps on Robertson: sensitivities outside the error test there
builder = builder.sensitivities() if (method == "tr_bdf2" and model.startswith("robertson")) else builder.sens_rtol(tol["rtol"]).sens_atol([1e-6] * nsa)
solver = getattr(builder.build(), method + "_sens")()
best = None
for it in range(3):
    ys, sens = solver.solve_dense_sensitivities(t_eval)
    if it > 0 and (best is None or solver.last_kernel_ms() < best):
        best = solver.last_kernel_ms()
st = solver.statistics_array()
print(json.dumps({"model": model, "method": method, "batch": B, "kernel_ms": best, "integrator_ms": solver.last_integrator_ms(), "instances_per_s": B / best * 1e3,
                  "newton_iters_per_s": float(st[:, 8].sum()) / best * 1e3, "steps_mean": float(st[:, 6].mean()),
                  "nli_mean": float(st[:, 8].mean()), "failed": int((solver.status() != 0).sum())}))
