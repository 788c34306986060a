#!/usr/bin/env python
"""bench.py -- BDF Newton-iterations/s over a batch of Robertson instances (BASELINE.json config 2).

One "step" = one `problem.bdf().solve_dense(t_eval)` pass over the whole batch (10^6 instances per
GPU, rate-constant sweep, t in [0, 1e4]) through the C ABI of libdiffsol_b200.so.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]

N > 1 is launched by torchrun (one rank per GPU, NCCL); instances shard across ranks with no
collective inside the time loop and ONE all-gather of the trajectories at the end of each step.
`--impl reference` times the CPU restatement of the reference path (oracle/, reference-literal libm
pow) on all host threads, on a bounded sample of the same workload.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "bdf_newton_iters_per_sec"
UNIT = "newton_iters/s"
WORKLOAD = "robertson_ode n=3 rate-constant sweep, Bdf, t in [0,1e4], 6 t_eval, rtol 1e-4 atol [1e-8,1e-14,1e-6]"


def algorithmic_bytes(stats_sum, n, npar, nt, nbatch, mass_words, history_cols=5 + 3):
    """bytes an HBM-resident implementation must move, from the counters.  `history_cols`: the columns
    of per-step history read and written by every attempted step -- the BDF difference array (max order + 3), or the s stage
    derivatives of an (E)SDIRK method."""
    b_nl = 8 * (n * n + 4 * n + npar) + 4 * n
    b_lu = 8 * (n * n + mass_words + n * n) + 4 * n
    b_j = 8 * (n * n + n + npar)
    b_st = 8 * (2 * history_cols * n + 3 * n)
    nli, setups, me = stats_sum["nli"], stats_sum["setups"], stats_sum["me"]
    attempts = stats_sum["steps"] + stats_sum["etf"] + stats_sum["nlf"]
    return nli * b_nl + setups * b_lu + me * b_j + attempts * b_st + nbatch * nt * 8 * n


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.f.close()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); smax.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # samples under load = the upper half (idle samples before/after the region drag the median down)
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons),
                "samples": len(sm)}


def run_reference(args):
    """The reference's own CPU path (restated in oracle/, libm pow = reference-literal), all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import oracle as orc
    from diffsol_b200 import sweeps
    orc.build()
    threads = orc.num_threads()
    desc = orc.make_desc("robertson_ode", powmode=0, **sweeps.ROBERTSON_ODE_TOL)
    sample = args.cpu_sample or (1 << 20)
    p = sweeps.robertson_sweep(np.arange(sample))
    for _ in range(max(args.warmup, 1)):
        orc.batch_solve_dense(desc, p[: max(256, sample // 16)], sweeps.ROBERTSON_T_EVAL, nthreads=threads)
    t0 = time.perf_counter()
    nli = 0
    for _ in range(args.steps):
        _, stats, status = orc.batch_solve_dense(desc, p, sweeps.ROBERTSON_T_EVAL, nthreads=threads)
        nli += int(stats[:, 8].sum())
    dt = time.perf_counter() - t0
    value = nli / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample_instances_per_step": sample, "host_threads": threads},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "%d instances/step x %d steps of the same sweep, oracle with libm pow" % (sample, args.steps)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "instances_per_sec": sample * args.steps / dt,
    }
    print(json.dumps(line), flush=True)
    return 0


def cpu_baseline(seconds_target=15.0):
    from oracle import oracle as orc
    from diffsol_b200 import sweeps
    orc.build()
    threads = orc.num_threads()
    desc = orc.make_desc("robertson_ode", powmode=0, **sweeps.ROBERTSON_ODE_TOL)
    probe = 1024
    p = sweeps.robertson_sweep(np.arange(probe))
    t0 = time.perf_counter()
    orc.batch_solve_dense(desc, p, sweeps.ROBERTSON_T_EVAL, nthreads=threads)
    rate = probe / (time.perf_counter() - t0)
    sample = int(min(max(rate * seconds_target, probe), 1 << 22))
    p = sweeps.robertson_sweep(np.arange(sample))
    t0 = time.perf_counter()
    _, stats, _ = orc.batch_solve_dense(desc, p, sweeps.ROBERTSON_T_EVAL, nthreads=threads)
    dt = time.perf_counter() - t0
    return {"value": float(stats[:, 8].sum() / dt), "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "first %d instances of the same sweep, %.1f s, oracle (CPU restatement, libm pow), %d threads"
                      % (sample, dt, threads),
            "instances_per_sec": sample / dt}


def band_algorithmic_bytes(st, n, npar, nt, nout, B, has_mass):
    """The SURVEY 8(d) formula with the band storage the banded kernels really read (kl = ku = 1): factors (2kl+ku+1) n
    + n pivots, Jacobian / mass (kl+ku+1) n each."""
    ldab, ldj = 4, 3
    nli, setups, me = int(st[:, 8].sum()), int(st[:, 0].sum()), int(st[:, 12].sum())
    attempts = int(st[:, 6].sum() + st[:, 7].sum() + st[:, 9].sum())
    mass = ldj * n if has_mass else 0
    return (nli * 8 * (ldab * n + n + 4 * n + npar) + setups * 8 * (ldj * n + mass + ldab * n + n)
            + me * 8 * (ldj * n + n + npar) + attempts * 8 * 19 * n + B * nt * 8 * nout)


def parity_sample(orc, desc, params, t_eval, ys, stats, status, rtol, atol, roots=False):
    """CUDA result of a sample of the sweep against the oracle on the same inputs: fraction of instances whose 13 integer
    counters and status agree, and the reference's own state measure sqrt(mean(((y - y*) / (|y*| rtol + atol))^2))
    (ode_solver/mod.rs:164-173; its acceptance bound is 20), maximised over instances and output times."""
    if roots:
        ys_o, stats_o, status_o = orc.batch_solve_dense_roots(desc, params, t_eval)[:3]
    else:
        ys_o, stats_o, status_o = orc.batch_solve_dense(desc, params, t_eval)
    same = (stats[:, :13] == stats_o[:, :13]).all(axis=1) & (status == status_o)
    ok = np.isfinite(ys_o) & np.isfinite(ys)
    at = np.broadcast_to(np.asarray(atol, dtype=np.float64), ys_o.shape[-1:]) if np.ndim(atol) else float(atol)
    w = np.abs(ys_o) * rtol + at
    e = np.where(ok, (ys - ys_o) / w, 0.0)
    err = float(np.sqrt((e * e).mean(axis=-1)).max()) if e.size else 0.0
    return {"sample_instances": int(len(params)), "counters_equal_frac": float(same.mean()),
            "states_bit_identical": bool(np.array_equal(ys, ys_o, equal_nan=True)), "max_weighted_state_err": err}


def time_config(solver, t_eval, reps=2):
    """min over reps of the whole device time of one device-resident solve_dense pass (every kernel between the C ABI's own
    events, parameters already on the device); the first call (workspace allocation) is a warm-up.  Returns (ms,
    integrator-kernel ms)."""
    import torch
    pr = solver.problem
    out = torch.empty((len(t_eval) * pr.nout, pr.nbatch), dtype=torch.float64, device=torch.device("cuda", torch.cuda.current_device()))
    solver.set_params()
    best, integ = None, None
    for it in range(reps + 1):
        solver.solve_dense_device(t_eval, out.data_ptr())
        torch.cuda.synchronize()
        ms = solver.last_kernel_ms()
        if it > 0 and (best is None or ms < best):
            best, integ = ms, solver.last_integrator_ms()
    return best, integ


def other_configs(diffsol_b200, sweeps, orc, rank, world, local_rank, peak):
    """BASELINE.json configs 3, 4 and 5 on THIS rank's shard (instances i = rank + world * k), outside the headline timed
    region: ms per pass, instances/s, Newton-it/s, the band-byte HBM fraction, and a parity sample against the oracle."""
    out = {}
    S = sweeps
    # ---- config 3: Van der Pol mu in [1, 1e6], B = 4e6, TR-BDF2 ---------------------------------------------------
    B = 4000000 // world
    gidx = rank + world * np.arange(B, dtype=np.int64)
    p = S.van_der_pol_scaled_sweep(gidx)
    prob = diffsol_b200.OdeBuilder().rhs_implicit("van_der_pol_scaled").p(p).rtol(1e-4).atol([1e-6]).device(local_rank).build()
    solver = prob.tr_bdf2()
    ms, integ = time_config(solver, S.VAN_DER_POL_T_EVAL)
    status = solver.status()
    st = solver.statistics_array()
    done = status == 0
    c3 = {"workload": "van_der_pol (scaled time) mu=10^(6u), TR-BDF2, 8 t_eval, rtol 1e-4 atol 1e-6", "instances": B, "ms": ms,
          "instances_per_s": B / ms * 1e3, "newton_iters_per_s": float(st[:, 8].sum()) / ms * 1e3,
          "failed_instances": int((~done).sum()),
          "failed_note": "TooManyNonlinearSolverFailures under the reference's default limit of 50 (runge_kutta.rs:869-884)",
          "completed_instances_per_s": float(done.sum()) / ms * 1e3,
          "newton_iters_per_s_of_completed_instances": float(st[done, 8].sum()) / ms * 1e3}
    # the SURVEY 8(d) byte model with TR-BDF2's three stage derivatives as the per-step history (dense n = 2 storage)
    ssum = {"nli": int(st[:, 8].sum()), "setups": int(st[:, 0].sum()), "me": int(st[:, 12].sum()), "steps": int(st[:, 6].sum()),
            "etf": int(st[:, 7].sum()), "nlf": int(st[:, 9].sum())}
    alg = algorithmic_bytes(ssum, 2, int(p.shape[1]), len(S.VAN_DER_POL_T_EVAL), B, 0, history_cols=3)
    c3["algorithmic_GB"] = alg / 1e9
    c3["frac_hbm"] = alg / (ms * 1e-3) / 1e9 / peak
    c3["kernel"] = "dsb_sdirk_solve_dense_kernel<ModelVanDerPolScaled> (thread per instance, state on chip)"
    nsamp = 4096
    ps = S.van_der_pol_scaled_sweep(np.arange(nsamp))
    ss = diffsol_b200.OdeBuilder().rhs_implicit("van_der_pol_scaled").p(ps).rtol(1e-4).atol([1e-6]).device(local_rank).build().tr_bdf2()
    ys = ss.solve_dense(S.VAN_DER_POL_T_EVAL)
    c3["parity"] = parity_sample(orc, orc.make_desc("van_der_pol_scaled", method="tr_bdf2", powmode=1, rtol=1e-4, atol=1e-6), ps,
                                 S.VAN_DER_POL_T_EVAL, ys, ss.statistics_array(), ss.status(), 1e-4, 1e-6)
    out["config3_van_der_pol_trbdf2"] = c3
    del solver, prob, ss
    # ---- config 4: heat-equation DAE n = 256, B = 16384, BDF ------------------------------------------------------
    def heat_params(idx):
        return np.stack([1.0 + S.uniform(idx, 0), 0.1 + 0.3 * S.uniform(idx, 1), 0.6 + 0.3 * S.uniform(idx, 2)], axis=1)
    B = 16384 // world
    gidx = rank + world * np.arange(B, dtype=np.int64)
    t_eval = np.arange(1, 101) / 100.0 * 0.99
    prob = diffsol_b200.OdeBuilder().rhs_implicit("heat1d_dae_256").p(heat_params(gidx)).rtol(1e-6).atol(1e-6).device(local_rank).build()
    solver = prob.bdf()
    ms, integ = time_config(solver, t_eval)
    st = solver.statistics_array()
    band = band_algorithmic_bytes(st, 256, 3, len(t_eval), 256, B, True)
    c4 = {"workload": "heat-equation DAE n=256 (singular mass), plateau IC sweep, Bdf, 100 t_eval, rtol=atol=1e-6", "instances": B,
          "ms": ms, "integrator_kernel_ms": integ, "kernel": "dsb_wband_bdf_solve_dense_kernel (warp per instance)",
          "instances_per_s": B / ms * 1e3, "newton_iters_per_s": float(st[:, 8].sum()) / ms * 1e3,
          "failed_instances": int((solver.status() != 0).sum()),
          "band_algorithmic_GB": band / 1e9, "band_frac_hbm": band / (ms * 1e-3) / 1e9 / peak}
    nsamp = 128
    ps = heat_params(np.arange(nsamp))
    ss = diffsol_b200.OdeBuilder().rhs_implicit("heat1d_dae_256").p(ps).rtol(1e-6).atol(1e-6).device(local_rank).build().bdf()
    ys = ss.solve_dense(t_eval)
    c4["parity"] = parity_sample(orc, orc.make_desc("heat1d_dae_256", powmode=1, rtol=1e-6, atol=1e-6), ps, t_eval, ys,
                                 ss.statistics_array(), ss.status(), 1e-6, 1e-6)
    out["config4_heat_dae_256"] = c4
    del solver, prob, ss
    # ---- config 5: battery model with its output (terminal voltage every 3 s) and stop (voltage cut-offs) functions ----
    for key, model, n, nsamp in (("config5_battery_n42_out_stop", "spm_stop", 42, 2048), ("config5_battery_n200_out_stop", "spm99_stop", 200, 128)):
        B = 2000000 // 8                    # BASELINE config 5: 2e6 instances over 8 GPUs = 250 000 per GPU
        gidx = rank + world * np.arange(B, dtype=np.int64)
        t_eval = np.arange(1, 1201) * 3.0
        cur = (0.6 + 0.8 * S.uniform(gidx, 0)).reshape(-1, 1)
        prob = diffsol_b200.OdeBuilder().rhs_implicit(model).p(cur).use_coloring(True).device(local_rank).build()
        solver = prob.bdf()
        ms, integ = time_config(solver, t_eval, reps=1)
        st = solver.statistics_array()
        band = band_algorithmic_bytes(st, n, 1, len(t_eval), 1, B, False)
        c5 = {"workload": "%s (SPM, n=%d, out = terminal voltage every 3 s to 3600 s, stop = voltage cut-offs), I=0.6+0.8u, Bdf, coloured J" % (model, n),
              "instances": B, "ms": ms, "integrator_kernel_ms": integ, "instances_per_s": B / ms * 1e3,
              "newton_iters_per_s": float(st[:, 8].sum()) / ms * 1e3, "failed_instances": int((solver.status() != 0).sum()),
              "stopped_on_voltage_cut_off": int((solver.root_info()[0] >= 0).sum()),
              "band_algorithmic_GB": band / 1e9, "band_frac_hbm": band / (ms * 1e-3) / 1e9 / peak}
        ps = (0.6 + 0.8 * S.uniform(np.arange(nsamp), 0)).reshape(-1, 1)
        ss = diffsol_b200.OdeBuilder().rhs_implicit(model).p(ps).use_coloring(True).device(local_rank).build().bdf()
        ys = ss.solve_dense(t_eval)
        c5["parity"] = parity_sample(orc, orc.make_desc(model, powmode=1, use_coloring=True), ps, t_eval, ys, ss.statistics_array(),
                                     ss.status(), 1e-6, 1e-6, roots=True)
        out[key] = c5
        del solver, prob, ss
    # ---- forward sensitivities: the headline workload with d y / d k1, d k2, d k3 beside the state ----
    import torch
    B = 1000000 // world if world > 1 else 1000000
    gidx = rank + world * np.arange(B, dtype=np.int64)
    tol = S.ROBERTSON_ODE_TOL
    def sens_solver(pp):
        return (diffsol_b200.OdeBuilder().rhs_implicit("robertson_ode").p(pp).rtol(tol["rtol"]).atol(tol["atol"])
                .sens_rtol(tol["rtol"]).sens_atol([1e-6] * 3).device(local_rank).build().bdf_sens())
    solver = sens_solver(S.robertson_sweep(gidx))
    t_eval = S.ROBERTSON_T_EVAL
    dev = torch.device("cuda", torch.cuda.current_device())
    ys_d = torch.empty((len(t_eval) * 3, B), dtype=torch.float64, device=dev)
    ss_d = torch.empty((len(t_eval) * 9, B), dtype=torch.float64, device=dev)
    solver.set_params()
    best = None
    for it in range(3):
        solver.solve_dense_sensitivities_device(t_eval, ys_d.data_ptr(), ss_d.data_ptr())
        torch.cuda.synchronize()
        if it > 0 and (best is None or solver.last_kernel_ms() < best):
            best = solver.last_kernel_ms()
    st = solver.statistics_array()
    cs = {"workload": "config 2 (robertson_ode sweep, Bdf) with forward sensitivities to the 3 rate constants in the error test "
                      "(sens_rtol = rtol = 1e-4, sens_atol 1e-6): problem.bdf_sens().solve_dense_sensitivities", "instances": B, "ms": best,
          "kernel": "dsb_bdf_solve_dense_kernel<DsbWithSens<ModelRobertsonOde<1>>> (thread per instance, state and 3 difference arrays on chip)",
          "instances_per_s": B / best * 1e3, "newton_iters_per_s": float(st[:, 8].sum()) / best * 1e3,
          "newton_iters_note": "state and sensitivity solves together: (1 + np) solves per attempted step on one factorisation",
          "failed_instances": int((solver.status() != 0).sum())}
    del solver, ys_d, ss_d
    nsamp = 1024
    ps = S.robertson_sweep(np.arange(nsamp))
    ssv = sens_solver(ps)
    ys, sens = ssv.solve_dense_sensitivities(t_eval)
    ys_o, se_o, st_o, status_o = orc.batch_solve_dense_sens(
        orc.make_desc("robertson_ode", powmode=1, sens=True, sens_rtol=tol["rtol"], sens_atol=[1e-6] * 3, **tol), ps, t_eval)
    cs["parity"] = {"sample_instances": nsamp,
                    "counters_and_status_equal": bool(np.array_equal(ssv.statistics_array()[:, :13], st_o[:, :13]) and np.array_equal(ssv.status(), status_o)),
                    "states_bitwise_equal": bool(np.array_equal(ys, ys_o, equal_nan=True)),
                    "sensitivities_bitwise_equal": bool(np.array_equal(sens, se_o, equal_nan=True))}
    out["sensitivities_robertson_bdf"] = cs
    del ssv
    return out


def fp64_peak():
    """FP64 micro-benchmark of THIS box (tools/fp64_peak.cu): the denominator SURVEY 8(d) asks for before any flop statement."""
    exe = os.path.join(ROOT, "tools", "_bin", "fp64_peak")
    try:
        if not os.path.exists(exe):
            os.makedirs(os.path.dirname(exe), exist_ok=True)
            subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "--fmad=false", "-std=c++17", "-o", exe,
                                   os.path.join(ROOT, "tools", "fp64_peak.cu")], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        return json.loads(subprocess.run([exe], stdout=subprocess.PIPE, text=True, timeout=120).stdout.strip().splitlines()[-1])
    except Exception as e:      # noqa: BLE001 -- the bench line must survive a missing nvcc
        return {"unavailable": str(e)[:120]}


def bind_to_gpu_numa_node(torch, local_rank):
    """Multi-GPU runs: restrict this rank to the CPUs NVML reports as local to its GPU BEFORE the pinned host buffers are
    allocated and first touched, so that the host <-> device copies of the end-to-end leg do not cross the socket
    interconnect.  Returns the number of CPUs the rank is bound to (0: left alone)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        props = torch.cuda.get_device_properties(local_rank)
        try:
            h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(props.uuid)).encode())
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = (max(os.cpu_count() or 64, 64) + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        local = {w * 64 + k for w, word in enumerate(mask) for k in range(64) if (int(word) >> k) & 1}
        cpus = sorted(local & os.sched_getaffinity(0))
        if not cpus or len(cpus) == len(os.sched_getaffinity(0)):
            return 0
        os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=1000000, help="instances per GPU")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip BASELINE configs 3-5, the parity samples and the FP64 probe")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import diffsol_b200
    from diffsol_b200 import capi, sweeps
    from diffsol_b200 import distributed as dsbd

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    capi.require_device()                       # fail loudly: there is no CPU fallback
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa_cpus = bind_to_gpu_numa_node(torch, local_rank) if world > 1 else 0
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    B = args.batch                               # weak scaling: per-GPU work is fixed
    n, npar = 3, 3
    t_eval = sweeps.ROBERTSON_T_EVAL
    nt = len(t_eval)
    gidx = rank + world * np.arange(B, dtype=np.int64)          # global instance ids of this shard
    params_host = torch.from_numpy(sweeps.robertson_sweep(gidx)).pin_memory()
    # tolerances of the reference's robertson_ode test problem
    problem = (diffsol_b200.OdeBuilder().rhs_implicit("robertson_ode").p(params_host.numpy())
               .rtol(sweeps.ROBERTSON_ODE_TOL["rtol"]).atol(sweeps.ROBERTSON_ODE_TOL["atol"]).device(local_rank).build())
    solver = problem.bdf()
    L = capi.lib()
    vp = ctypes.c_void_p

    params_dev = params_host.to(dev)
    ys_dev = torch.empty((nt * n, B), dtype=torch.float64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2
    stream = torch.cuda.current_stream()
    te = np.ascontiguousarray(t_eval)
    launches = [0]

    def step_device():
        capi.check(L.dsb_batch_solve_dense(solver._b, 0, vp(te.ctypes.data), nt, vp(ys_dev.data_ptr()), vp(stream.cuda_stream)))
        launches[0] += solver.last_launch_count()
        if world > 1:
            return dsbd.all_gather_batch_major(ys_dev, B * world)
        return ys_dev

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    capi.check(L.dsb_batch_set_params_device(solver._b, vp(params_dev.data_ptr()), B, npar, vp(stream.cuda_stream)))
    for _ in range(args.warmup):
        flush.zero_()
        step_device()
    barrier()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # ---- timed region: K steps, device time, L2 flushed between steps (flush excluded via events) ----
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches[0] = 0
    integ_ms = []
    barrier()
    for k in range(args.steps):
        flush.zero_()
        ev[k][0].record(stream)
        step_device()
        ev[k][1].record(stream)
        torch.cuda.synchronize()
        integ_ms.append(solver.last_integrator_ms())
    barrier()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = float(sum(step_ms))
    clocks = sampler.stop() if rank == 0 else None

    stats_names = {"nli": "number_of_nonlinear_solver_iterations", "setups": "number_of_linear_solver_setups",
                   "me": "rhs_number_of_matrix_evals", "steps": "number_of_steps",
                   "etf": "number_of_error_test_failures", "nlf": "number_of_nonlinear_solver_fails"}
    ssum = {k: solver.sum_statistic(v) for k, v in stats_names.items()}
    n_failed = int((solver.status() != 0).sum())

    # ---- end-to-end through the public API with HOST buffers (pinned), copies inside the timed region ----
    # One call = what `problem.bdf().solve_dense(t_eval)` hands back in the reference: the trajectories (plus the
    # per-instance status that stands for its Result<>).  Statistics are a separate getter there
    # (`get_statistics()`), so they are read after the timed region, through the device reduction.
    ys_host = torch.empty((B, nt, n), dtype=torch.float64).pin_memory()
    status_host = torch.empty((B,), dtype=torch.int32).pin_memory()

    def step_e2e():
        capi.check(L.dsb_batch_solve_dense_host(
            solver._b, 0, vp(params_host.data_ptr()), npar, vp(te.ctypes.data), nt,
            vp(ys_host.data_ptr()), None, vp(status_host.data_ptr())))

    step_e2e()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(2, min(args.steps, 3))
    for _ in range(e2e_steps):
        step_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0
    e2e_nli = solver.sum_statistic("number_of_nonlinear_solver_iterations") * e2e_steps
    assert int((status_host != 0).sum()) == n_failed and (n_failed > 0 or bool(torch.isfinite(ys_host).all()))

    t_ms = torch.tensor([total_ms, e2e_s * 1e3, float(np.mean(integ_ms))], dtype=torch.float64, device=dev)
    cnt = torch.tensor([ssum["nli"], ssum["setups"], ssum["me"], ssum["steps"], ssum["etf"], ssum["nlf"],
                        e2e_nli, n_failed, launches[0]], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    total_ms, e2e_ms, integ_mean_ms = [float(x) for x in t_ms.tolist()]
    c = [int(x) for x in cnt.tolist()]
    nli_all = c[0]

    # ---- outside the timed region: parity samples, the other BASELINE configs, the FP64 probe, the multi-GPU gather check ----
    extras = {}
    if not args.no_configs:
        from oracle import oracle as orc
        orc.build()
        peak_hbm = 6650.0
        try:
            peak_hbm = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0))
        except OSError:
            pass
        if world > 1:
            # the gathered block equals a single-GPU recomputation, bitwise -- on a sample of every rank's shard
            gathered = step_device()                        # [nt * n, B * world], global instance order
            torch.cuda.synchronize()
            if rank == 0:
                samp = np.unique(np.concatenate([np.arange(r, B * world, world)[:: max(1, B // 32)][:32] for r in range(world)]))
                ps = sweeps.robertson_sweep(samp)
                ss = (diffsol_b200.OdeBuilder().rhs_implicit("robertson_ode").p(ps).rtol(sweeps.ROBERTSON_ODE_TOL["rtol"])
                      .atol(sweeps.ROBERTSON_ODE_TOL["atol"]).device(local_rank).build().bdf())
                ys_s = ss.solve_dense(t_eval)                                               # [S, nt, n]
                got = gathered[:, torch.from_numpy(samp).to(dev)].cpu().numpy().reshape(nt, n, len(samp)).transpose(2, 0, 1)
                extras["gathered_equals_recomputed"] = {"sample_instances": int(len(samp)), "ranks_covered": world,
                                                        "bitwise_equal": bool(np.array_equal(got, ys_s))}
        if rank == 0:
            # config 2 parity: a sample of the same sweep against the oracle with this repo's pow (bit-exact contract) and
            # with libm pow (the reference's own arithmetic)
            nsamp = 4096
            ps = sweeps.robertson_sweep(np.arange(nsamp))
            ss = (diffsol_b200.OdeBuilder().rhs_implicit("robertson_ode").p(ps).rtol(sweeps.ROBERTSON_ODE_TOL["rtol"])
                  .atol(sweeps.ROBERTSON_ODE_TOL["atol"]).device(local_rank).build().bdf())
            ys_s = ss.solve_dense(t_eval)
            par = {}
            for name, mode in (("vs_oracle_dsb_pow", 1), ("vs_oracle_libm_pow", 0)):
                desc = orc.make_desc("robertson_ode", powmode=mode, **sweeps.ROBERTSON_ODE_TOL)
                par[name] = parity_sample(orc, desc, ps, t_eval, ys_s, ss.statistics_array(), ss.status(),
                                          sweeps.ROBERTSON_ODE_TOL["rtol"], sweeps.ROBERTSON_ODE_TOL["atol"])
            extras["parity"] = par
            extras["fp64"] = fp64_peak()
            del ss
        del params_dev, ys_dev, flush, ys_host
        torch.cuda.empty_cache()
        cfg = other_configs(diffsol_b200, sweeps, orc, rank, world, local_rank, peak_hbm)
        if world > 1:
            # per-config aggregate over the ranks: max of the times, sum of the counts (each rank ran its own shard)
            keys = sorted(cfg)
            t = torch.tensor([cfg[k]["ms"] for k in keys] + [float(cfg[k]["instances"]) for k in keys]
                             + [cfg[k]["newton_iters_per_s"] * cfg[k]["ms"] for k in keys], dtype=torch.float64, device=dev)
            tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
            nk = len(keys)
            for i, k in enumerate(keys):
                ms_all = float(tmax[i])
                cfg[k]["all_ranks"] = {"ms_max_over_ranks": ms_all, "instances_total": int(tsum[nk + i]),
                                       "instances_per_s": float(tsum[nk + i]) / ms_all * 1e3,
                                       "newton_iters_per_s": float(tsum[2 * nk + i]) / ms_all}
        extras["configs"] = cfg

    if rank == 0:
        value = nli_all * args.steps / (total_ms * 1e-3)
        ssum_all = dict(nli=c[0], setups=c[1], me=c[2], steps=c[3], etf=c[4], nlf=c[5])
        # roofline of the integrator kernel on ONE GPU: algorithmic bytes of this rank's launch / its duration
        alg = algorithmic_bytes(ssum, n, npar, nt, B, 0)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        achieved = alg / (integ_mean_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "instances_per_gpu": B, "instances_total": B * world,
                       "parallelism": "instances sharded i mod G, one all-gather of trajectories" if world > 1 else "single GPU",
                       "l2": "256 MiB buffer rewritten between timed steps", "failed_instances": c[7],
                       "rank0_bound_to_gpu_local_cpus": numa_cpus},
            "instances_per_sec": B * world * args.steps / (total_ms * 1e-3),
            "newton_iters_per_step": nli_all,
            "e2e": {"value": c[6] / (e2e_ms * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": int(B * npar * 8 + nt * 8),
                    "d2h_bytes_per_step": int(B * nt * n * 8 + B * 4),
                    "ms_per_step": e2e_ms / e2e_steps, "steps": e2e_steps},
            "gpu_launches": c[8],
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         # dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of this kernel
                         # (profiles/r2_bdf_kernel_ncu_summary.txt: 149.5 + 179.3 MB for a 10^6-instance launch),
                         # per instance: mostly the trajectories, counters and status it writes
                         "traffic": 328.8 * B, "traffic_source": "ncu capture profiles/r2_bdf_kernel_ncu_summary.txt (328.8 B/instance), scaled to this launch",
                         "kernel": "dsb_bdf_solve_dense_kernel<ModelRobertsonOde<1>>",
                         "kernel_ms": integ_mean_ms, "algorithmic_bytes_per_launch": alg,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                         "note": "achieved = algorithmic bytes of an HBM-resident implementation (SURVEY 8d) / kernel time; this kernel keeps the state in registers/shared memory (see traffic) and is bound by FP64 dependency latency and lane divergence"},
            "clocks": clocks,
            "counters": ssum_all,
        }
        line.update(extras)
        if "fp64" in extras and "dfma_tflops" in extras["fp64"]:
            # what really bounds the headline kernel: FP64 issue x lane occupancy (static figures from the committed ncu
            # capture of this kernel, profiles/r2_*; the peak is measured live on this box)
            line["fp64"]["kernel_note"] = ("dsb_bdf_solve_dense_kernel: see profiles/ for sm__pipe_fp64_cycles_active and "
                                           "thread_inst_executed_per_inst_executed (lanes per instruction) of the latest capture")
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline()
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
