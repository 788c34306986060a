#!/usr/bin/env python
"""What would FP64 tensor-core (DMMA) trailing updates do to the controller?  DMMA computes d = a * b + c with ONE rounding;
the reference (nalgebra's LU) rounds the product and the sum separately.  The oracle is run twice on the same sweeps --
reference arithmetic, and with fused multiply-adds in the trailing updates of every LU factorisation -- and the outcomes are
compared: instances whose 13 integer counters differ, and the largest weighted state difference.  (CPU experiment; the
speed side of the question is tools/fp64_peak.cu: dmma_tflops vs dfma_tflops.)
   python tests/diagnostics/dmma_rounding_check.py > profiles/r2_dmma_rounding_check.json"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402
from diffsol_b200 import sweeps  # noqa: E402

orc.build()
L = orc.lib()


def heat_params(idx):
    return np.stack([1.0 + sweeps.uniform(idx, 0), 0.1 + 0.3 * sweeps.uniform(idx, 1), 0.6 + 0.3 * sweeps.uniform(idx, 2)], axis=1)


cases = [
    ("robertson_ode n=3 (config 2)", "robertson_ode", sweeps.robertson_sweep(np.arange(20000)), sweeps.ROBERTSON_T_EVAL, dict(sweeps.ROBERTSON_ODE_TOL)),
    ("robertson_ode_g3 n=9", "robertson_ode_g3", sweeps.robertson_sweep(np.arange(4000)), sweeps.ROBERTSON_T_EVAL, dict(rtol=1e-4, atol=1e-8)),
    ("heat1d_dae_32 n=32 (dense LU)", "heat1d_dae_32", heat_params(np.arange(2000)), np.arange(1, 101) / 100.0 * 0.99, dict(rtol=1e-6, atol=1e-6)),
    ("heat1d_dae_256 n=256 (config 4 through the DENSE LU: the n >= 64 panel case)", "heat1d_dae_256", heat_params(np.arange(96)), np.arange(1, 101) / 100.0 * 0.99, dict(rtol=1e-6, atol=1e-6)),
]
out = {}
for label, model, p, t_eval, tol in cases:
    desc = orc.make_desc(model, powmode=1, **tol)
    L.orc_set_fused_lu_updates(0)
    y0, s0, st0 = orc.batch_solve_dense(desc, p, t_eval)
    L.orc_set_fused_lu_updates(1)
    y1, s1, st1 = orc.batch_solve_dense(desc, p, t_eval)
    L.orc_set_fused_lu_updates(0)
    differ = (s0[:, :13] != s1[:, :13]).any(axis=1) | (st0 != st1)
    at = np.asarray(tol["atol"], dtype=np.float64)
    w = np.abs(y0) * tol["rtol"] + at
    ok = np.isfinite(y0) & np.isfinite(y1)
    out[label] = {"instances": int(len(p)), "instances_with_different_counters": int(differ.sum()), "fraction": float(differ.mean()),
                  "states_bit_identical_instances": int((np.where(ok, y0 == y1, True)).reshape(len(p), -1).all(axis=1).sum()),
                  "max_state_difference_in_tolerances": float(np.where(ok, np.abs(y1 - y0) / w, 0.0).max()),
                  "steps_total_reference_vs_fused": [int(s0[:, 6].sum()), int(s1[:, 6].sum())],
                  "newton_iterations_total_reference_vs_fused": [int(s0[:, 8].sum()), int(s1[:, 8].sum())]}
print(json.dumps(out, indent=1))
