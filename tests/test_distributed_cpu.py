"""world_size-2 gloo tests of the multi-GPU plumbing (CPU only): the i mod G partition, the single
all-gather of batch-major trajectories back into global instance order, and the ragged-tail padding.
The per-shard 'compute' here is the CPU oracle (tests may use it); on the GPU box bench.py runs the same
gather over NCCL with the sm_100a kernels producing the shards."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nbatch, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from diffsol_b200 import distributed as dsbd, sweeps
    from oracle import oracle as orc
    idx = sweeps.shard_indices(nbatch, rank, world)
    assert len(idx) == dsbd.local_count(nbatch, rank, world)
    p = sweeps.robertson_sweep(idx)
    desc = orc.make_desc("robertson_ode", powmode=1, **sweeps.ROBERTSON_ODE_TOL)
    t_eval = sweeps.ROBERTSON_T_EVAL[:3]
    ys, stats, status = orc.batch_solve_dense(desc, p, t_eval, nthreads=1)       # [B_r, nt, n]
    # device layout of the product: batch-major [nt * n, B_r]
    local = torch.from_numpy(np.ascontiguousarray(ys.reshape(len(idx), -1).T))
    gathered = dsbd.all_gather_batch_major(local, nbatch)
    st_local = torch.from_numpy(np.ascontiguousarray(stats[:, :13].T))
    st_gathered = dsbd.all_gather_batch_major(st_local, nbatch)
    np.save(os.path.join(out_dir, "ys_%d.npy" % rank), gathered.numpy())
    np.save(os.path.join(out_dir, "st_%d.npy" % rank), st_gathered.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("nbatch", [64, 37])        # 37: ragged tail (ranks own 19 and 18 instances)
def test_two_rank_gather_equals_single_rank(tmp_path, nbatch, oracle):
    from diffsol_b200 import sweeps
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, nbatch, str(tmp_path)), nprocs=world, join=True)
    desc = oracle.make_desc("robertson_ode", powmode=1, **sweeps.ROBERTSON_ODE_TOL)
    p = sweeps.robertson_sweep(np.arange(nbatch))
    ys, stats, status = oracle.batch_solve_dense(desc, p, sweeps.ROBERTSON_T_EVAL[:3], nthreads=2)
    want = ys.reshape(nbatch, -1).T
    for r in range(world):
        got = np.load(os.path.join(str(tmp_path), "ys_%d.npy" % r))
        assert got.shape == want.shape
        assert np.array_equal(got, want)                  # bitwise: every rank holds the global result
        st = np.load(os.path.join(str(tmp_path), "st_%d.npy" % r))
        assert np.array_equal(st, stats[:, :13].T)


def test_partition_covers_every_instance_once():
    from diffsol_b200 import distributed as dsbd, sweeps
    for nbatch in (1, 7, 8, 1000003):
        for world in (1, 2, 4, 8):
            counts = [dsbd.local_count(nbatch, r, world) for r in range(world)]
            assert sum(counts) == nbatch and max(counts) - min(counts) <= 1
    idx = np.concatenate([sweeps.shard_indices(1001, r, 8) for r in range(8)])
    assert np.array_equal(np.sort(idx), np.arange(1001))
    # the counter RNG is a pure function of the global index: shards regenerate their own parameters
    p_all = sweeps.robertson_sweep(np.arange(1001))
    assert np.array_equal(sweeps.robertson_sweep(sweeps.shard_indices(1001, 3, 8)), p_all[3::8])


def test_c_abi_library_exports_every_declared_symbol():
    """No compute calls without a GPU: the library loads, every symbol of include/diffsol_b200.h resolves,
    host-only entry points work, and a solve without a device fails loudly (no CPU fallback)."""
    import re
    import ctypes
    sys.path.insert(0, ROOT)
    import diffsol_b200
    from diffsol_b200 import capi
    L = capi.lib()
    header = open(os.path.join(ROOT, "include", "diffsol_b200.h")).read()
    declared = set(re.findall(r"\b(dsb_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    missing = [name for name in sorted(declared) if not hasattr(L, name)]
    assert not missing, missing
    assert set(capi.SIGNATURES) == declared
    o = capi.Options()
    L.dsb_options_default(ctypes.byref(o))
    assert o.max_nonlinear_solver_iterations == 10 and o.max_error_test_failures == 40
    assert o.nonlinear_solver_tolerance == 0.2 and o.threshold_to_update_jacobian == 0.3
    prob = diffsol_b200.OdeBuilder().rhs_implicit("robertson_dae").p([0.04, 1e4, 3e7]).build()
    assert (prob.nstates, prob.nparams, prob.has_mass) == (3, 3, True)
    if capi.device_count() == 0:
        with pytest.raises(diffsol_b200.DiffsolB200Error):
            prob.bdf()


def test_c_abi_argument_errors_are_reported_not_swallowed():
    """Host-only entry points of the C ABI: bad arguments come back as DSB_BAD_ARG (-2) with a message in the thread-local
    dsb_last_error(), the conventions of diffsol-c (c_api_utils.rs:3-5, error_c.rs:12-46); no GPU needed."""
    import ctypes
    sys.path.insert(0, ROOT)
    from diffsol_b200 import capi
    L = capi.lib()
    h = ctypes.c_void_p()
    assert L.dsb_problem_new(10_000, ctypes.byref(h)) == -2 and b"unknown model" in L.dsb_last_error()
    assert L.dsb_problem_new(capi.MODELS["robertson_dae"], None) == -2
    assert L.dsb_problem_new(capi.MODELS["robertson_dae"], ctypes.byref(h)) == 0
    two = (ctypes.c_double * 2)(1e-6, 1e-6)
    assert L.dsb_problem_set_atol(h, two, 2) == -2 and b"atol" in L.dsb_last_error()     # 1 or nstates entries
    three = (ctypes.c_double * 3)(1e-8, 1e-6, 1e-6)
    assert L.dsb_problem_set_atol(h, three, 3) == 0 and L.dsb_problem_set_atol(h, two, 1) == 0
    assert L.dsb_problem_set_h0(h, ctypes.c_double(0.0)) == -2 and L.dsb_problem_set_h0(h, ctypes.c_double(-1.0)) == 0
    nout = ctypes.c_int32(-1)
    assert L.dsb_problem_nout(h, ctypes.byref(nout)) == 0 and nout.value == 3             # no output function: the states
    assert L.dsb_problem_free(h) == 0
    h2 = ctypes.c_void_p()
    assert L.dsb_problem_new(capi.MODELS["spm_stop"], ctypes.byref(h2)) == 0
    assert L.dsb_problem_nout(h2, ctypes.byref(nout)) == 0 and nout.value == 1            # terminal voltage
    assert L.dsb_problem_free(h2) == 0
    assert L.dsb_version().decode()
