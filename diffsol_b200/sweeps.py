"""Synthetic parameter / initial-condition sweeps for the BASELINE.json configurations.

Counter-based so that the CPU oracle and every GPU rank generate the identical instance i without
communicating: u_j(i) = splitmix64(0xD1FF501 ^ (8 i + j)) / 2^64   (SURVEY.md section 8d).
"""
import numpy as np

_MASK = (1 << 64) - 1


def splitmix64(x):
    """Vectorised splitmix64 finaliser over uint64 arrays."""
    x = np.asarray(x, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = x + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def uniform(i, j):
    """u_j(i) in [0, 1) for instance indices i (array) and stream j."""
    i = np.asarray(i, dtype=np.uint64)
    key = np.uint64(0xD1FF501) ^ (i * np.uint64(8) + np.uint64(j))
    return (splitmix64(key) >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def robertson_sweep(indices):
    """Config 2: rate constants k1 = 0.04 * 10^(u0 - 1/2), k2 = 1e4 * 10^(u1 - 1/2), k3 = 3e7 * 10^(u2 - 1/2).
    -> params[len(indices), 3], instance-major."""
    idx = np.asarray(indices, dtype=np.uint64)
    base = np.array([0.04, 1.0e4, 3.0e7])
    p = np.empty((len(idx), 3))
    for j in range(3):
        p[:, j] = base[j] * 10.0 ** (uniform(idx, j) - 0.5)
    return p


ROBERTSON_T_EVAL = np.array([0.4, 4.0, 40.0, 400.0, 4000.0, 1.0e4])
ROBERTSON_ODE_TOL = dict(rtol=1e-4, atol=[1e-8, 1e-14, 1e-6])     # robertson_ode.rs:55-65
ROBERTSON_DAE_TOL = dict(rtol=1e-4, atol=[1e-8, 1e-6, 1e-6])      # robertson.rs:103-105


def van_der_pol_sweep(indices):
    """Config 3: mu = 10^(6 u0) -> params[len, 1]."""
    idx = np.asarray(indices, dtype=np.uint64)
    return (10.0 ** (6.0 * uniform(idx, 0))).reshape(-1, 1)


def van_der_pol_scaled_sweep(indices):
    """Config 3 with per-instance end time T = max(20, 2 mu) folded into the equations (model
    van_der_pol_scaled, scaled time tau = t / T in [0, 1]) -> params[len, 2] = [mu, T]."""
    mu = van_der_pol_sweep(indices)[:, 0]
    return np.stack([mu, np.maximum(20.0, 2.0 * mu)], axis=1)


VAN_DER_POL_T_EVAL = np.arange(1, 9) / 8.0            # 8 equally spaced points in scaled time
VAN_DER_POL_TOL = dict(rtol=1e-4, atol=[1e-6])


def shard_indices(nbatch, rank, world):
    """Instance i -> rank i mod world (interleaved keeps a parameter-sorted sweep balanced)."""
    return np.arange(rank, nbatch, world, dtype=np.int64)
