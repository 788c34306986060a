#!/usr/bin/env python
"""Time one BASELINE.json configuration other than the headline one (which bench.py owns) on one GPU:
   python tools/bench_config.py heat256 [batch] [auto|block|band] [bdf|tr_bdf2|esdirk34]  |  spm | spm99 | spm_stop | spm99_stop [batch] [bdf|tr_bdf2|esdirk34]  |  vdp | robertson_dae [batch]
Prints ms per pass, instances/s, Newton-it/s and the algorithmic-byte HBM roofline fraction."""
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
which = sys.argv[1]
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
if which.startswith("heat"):
    n = int(which[4:])
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
    idx = np.arange(B)
    p = np.stack([1.0 + sweeps.uniform(idx, 0), 0.1 + 0.3 * sweeps.uniform(idx, 1), 0.6 + 0.3 * sweeps.uniform(idx, 2)], axis=1)
    t_eval = np.arange(1, 101) / 100.0 * 0.99
    prob = ds.OdeBuilder().rhs_implicit("heat1d_dae_%d" % n).p(p).rtol(1e-6).atol(1e-6).build()
    solver, npar, mass_words = getattr(prob, sys.argv[4] if len(sys.argv) > 4 else "bdf")(), 3, n * n
    if len(sys.argv) > 3:
        solver.set_execution(sys.argv[3])
elif which in ("spm", "spm99", "spm_stop", "spm99_stop"):
    n, npar, mass_words = (42 if which in ("spm", "spm_stop") else 200), 1, 0
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 250000
    p = (0.6 + 0.8 * sweeps.uniform(np.arange(B), 0)).reshape(-1, 1)
    # states-only models: 12 output points; the full model (output = terminal voltage, stop = voltage cut-offs): the
    # reference example's output every 3 s over 3600 s (examples/physics-based-battery-simulation/src/main.rs:20-21)
    t_eval = np.arange(1, 1201) * 3.0 if which.endswith("_stop") else np.arange(1, 13) * 300.0
    prob = ds.OdeBuilder().rhs_implicit(which).p(p).use_coloring(True).build()
    solver = getattr(prob, sys.argv[3] if len(sys.argv) > 3 else "bdf")()
elif which == "vdp":
    n, npar, mass_words = 2, 2, 0
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 4000000
    p = sweeps.van_der_pol_scaled_sweep(np.arange(B))
    t_eval = sweeps.VAN_DER_POL_T_EVAL
    prob = ds.OdeBuilder().rhs_implicit("van_der_pol_scaled").p(p).rtol(1e-4).atol([1e-6]).build()
    solver = prob.tr_bdf2()
elif which == "robertson_dae":
    n, npar, mass_words = 3, 3, 3
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 1000000
    p = sweeps.robertson_sweep(np.arange(B))
    t_eval = sweeps.ROBERTSON_T_EVAL
    prob = ds.OdeBuilder().rhs_implicit("robertson_dae").p(p).rtol(1e-4).atol([1e-8, 1e-6, 1e-6]).build()
    solver = prob.bdf()
else:
    raise SystemExit("unknown configuration")

times = []
for it in range(3):
    t0 = time.perf_counter()
    ys = solver.solve_dense(t_eval)
    times.append((time.perf_counter() - t0, solver.last_kernel_ms()))
integ_ms = solver.last_integrator_ms()
st = solver.statistics_array()
status = solver.status()
kms = min(k for _, k in times[1:])
nli, setups, me = int(st[:, 8].sum()), int(st[:, 0].sum()), int(st[:, 12].sum())
attempts = int(st[:, 6].sum() + st[:, 7].sum() + st[:, 9].sum())
nt = len(t_eval)
alg = (nli * (8 * (n * n + 4 * n + npar) + 4 * n) + setups * (8 * (2 * n * n + mass_words) + 4 * n)
       + me * 8 * (n * n + n + npar) + attempts * 8 * 19 * n + B * nt * 8 * prob.nout)
# banded path (dsb_band_bdf_kernel.cuh): the same formula with the band storage it really reads (kl = ku = 1):
# factors (2kl+ku+1) n + n pivots, Jacobian (kl+ku+1) n
band = None
if which.startswith("spm") or which.startswith("heat"):
    ldab, ldj = 4, 3
    band_mass = ldj * n if which.startswith("heat") else 0
    band = (nli * 8 * (ldab * n + n + 4 * n + npar) + setups * 8 * (ldj * n + band_mass + ldab * n + n)
            + me * 8 * (ldj * n + n + npar) + attempts * 8 * 19 * n + B * nt * 8 * prob.nout)
print(json.dumps({"config": which, "n": n, "batch": B, "kernel_ms": kms, "integrator_ms": integ_ms, "e2e_ms": min(t for t, _ in times[1:]) * 1e3,
                  "instances_per_s": B / kms * 1e3, "newton_iters_per_s": nli / kms * 1e3,
                  "steps_mean": float(st[:, 6].mean()), "nli_mean": float(st[:, 8].mean()), "setups_mean": float(st[:, 0].mean()),
                  "failed": int((status != 0).sum()), "stopped_on_root": int((solver.root_info()[0] >= 0).sum()), "algorithmic_GB": alg / 1e9,
                  "achieved_GBps": alg / kms / 1e6, "frac_hbm": alg / kms / 1e6 / peak,
                  "band_algorithmic_GB": None if band is None else band / 1e9,
                  "band_frac_hbm": None if band is None else band / kms / 1e6 / peak}))
