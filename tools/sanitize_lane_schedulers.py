import os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
import diffsol_b200 as ds
from diffsol_b200 import sweeps
idx = np.arange(96)
p = sweeps.robertson_sweep(idx)
tol = sweeps.ROBERTSON_ODE_TOL
prob = ds.OdeBuilder().rhs_implicit("robertson_ode").p(p).rtol(tol["rtol"]).atol(tol["atol"]).build()
for m in ("bdf", "tr_bdf2"):
    s = getattr(prob, m)(); s.solve_dense(sweeps.ROBERTSON_T_EVAL[:4]); print(m, "ok", int((s.status() != 0).sum()))
s = (ds.OdeBuilder().rhs_implicit("robertson_ode").p(p).rtol(tol["rtol"]).atol(tol["atol"]).sens_rtol(tol["rtol"]).sens_atol([1e-6] * 3).build().bdf_sens())
s.solve_dense_sensitivities(sweeps.ROBERTSON_T_EVAL[:4]); print("bdf_sens ok")
cur = (0.6 + 0.8 * sweeps.uniform(np.arange(40), 0)).reshape(-1, 1)
s = ds.OdeBuilder().rhs_implicit("spm_stop").p(cur).use_coloring(True).build().bdf()
s.solve_dense(np.arange(1, 41) * 3.0); print("spm_stop band lane ok")
