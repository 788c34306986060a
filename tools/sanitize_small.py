"""Small runs of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import diffsol_b200 as ds
from diffsol_b200 import sweeps
idx = np.arange(200)
p = sweeps.robertson_sweep(idx)
for model, tol in (("robertson_ode", sweeps.ROBERTSON_ODE_TOL), ("robertson_dae", sweeps.ROBERTSON_DAE_TOL)):
    prob = ds.OdeBuilder().rhs_implicit(model).p(p).rtol(tol["rtol"]).atol(tol["atol"]).build()
    for m in ("bdf", "tr_bdf2", "esdirk34"):
        s = getattr(prob, m)(); s.solve_dense(sweeps.ROBERTSON_T_EVAL[:4]); print(model, m, "ok", int((s.status() != 0).sum()))
    s = prob.bdf().set_execution("block"); s.solve_dense(sweeps.ROBERTSON_T_EVAL[:3]); print(model, "block ok")
hp = np.stack([1.0 + sweeps.uniform(idx[:6], 0), 0.1 + 0.3 * sweeps.uniform(idx[:6], 1), 0.6 + 0.3 * sweeps.uniform(idx[:6], 2)], axis=1)
for model in ("heat1d_dae_32", "heat1d_dae_256"):
    s = ds.OdeBuilder().rhs_implicit(model).p(hp).rtol(1e-6).atol(1e-6).build().bdf()
    s.solve_dense(np.arange(1, 11) / 100.0); print(model, "ok", s.get_statistics(0)["number_of_steps"])
os.environ["DSB_COOP_DENSE_ONLY"] = "1"
s = ds.OdeBuilder().rhs_implicit("heat1d_dae_256").p(hp[:2]).rtol(1e-6).atol(1e-6).build().bdf()
s.solve_dense([0.01, 0.02]); print("heat256 dense ok")
