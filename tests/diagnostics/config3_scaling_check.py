#!/usr/bin/env python
"""BASELINE config 3 (Van der Pol, mu = 10^(6u), TR-BDF2, t_final = max(20, 2 mu)): is the high rate of
TooManyNonlinearSolverFailures an artefact of integrating in scaled time tau = t / T (model van_der_pol_scaled, which lets
one t_eval grid serve every instance)?  The same instances are integrated by the oracle in PHYSICAL time with their own
t_final (model van_der_pol, p = [mu], one solve per instance) and the per-instance outcomes are compared.
   python tests/diagnostics/config3_scaling_check.py [sample] > profiles/r2_config3_scaling_check.json"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402
from diffsol_b200 import sweeps  # noqa: E402

orc.build()
S = int(sys.argv[1]) if len(sys.argv) > 1 else 512
idx = np.arange(S)
ps = sweeps.van_der_pol_scaled_sweep(idx)
tau = sweeps.VAN_DER_POL_T_EVAL
out = {}
for limit in (50, 1000000):
    opts = dict(max_nonlinear_solver_failures=limit)
    ds = orc.make_desc("van_der_pol_scaled", method="tr_bdf2", powmode=1, rtol=1e-4, atol=1e-6, options=opts)
    _, st_s, status_s = orc.batch_solve_dense(ds, ps, tau)
    du = orc.make_desc("van_der_pol", method="tr_bdf2", powmode=1, rtol=1e-4, atol=1e-6, options=opts)
    status_u = np.zeros(S, dtype=np.int32)
    nli_u = np.zeros(S, dtype=np.int64)
    for k in range(S):
        _, st, stt = orc.batch_solve_dense(du, ps[k:k + 1, :1], tau * ps[k, 1])
        status_u[k] = stt[0]; nli_u[k] = st[0, 8]
    out["max_nonlinear_solver_failures=%d" % limit] = {
        "sample": S,
        "scaled_time": {"failed": int((status_s != 0).sum()), "status_histogram": {int(k): int(v) for k, v in zip(*np.unique(status_s, return_counts=True))},
                        "newton_iters_mean": float(st_s[:, 8].mean())},
        "physical_time": {"failed": int((status_u != 0).sum()), "status_histogram": {int(k): int(v) for k, v in zip(*np.unique(status_u, return_counts=True))},
                          "newton_iters_mean": float(nli_u.mean())},
        "same_outcome_frac": float(((status_s != 0) == (status_u != 0)).mean()),
        "smallest_mu_that_fails": {"scaled": float(ps[status_s != 0, 0].min()) if (status_s != 0).any() else None,
                                   "physical": float(ps[status_u != 0, 0].min()) if (status_u != 0).any() else None},
    }
print(json.dumps(out, indent=1))
