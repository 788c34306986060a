"""diffsol_b200 -- B200-native batched implicit ODE/DAE integration (BDF / SDIRK step loop of diffsol)."""
from .capi import DiffsolB200Error, MODELS, METHODS, STAT_NAMES, STATUS_NAMES, Options  # noqa: F401
from .ode import OdeBuilder, OdeSolverProblem, BatchedSolver  # noqa: F401
