#!/usr/bin/env python
"""BASELINE config 5 as it is specified: the battery model over the GPUs of one node, the batch sharded i mod G, ONE
all-gather of the output trajectories over NCCL.

    python tools/bench_config5_multi.py [--model spm99_stop] [--batch 250000] [--steps 3] [--warmup 1]      (1 GPU)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        tools/bench_config5_multi.py --batch 250000                                                           (8 GPUs)

--batch is PER GPU (weak scaling: 8 x 250000 = 2e6 instances, the configuration's size).  The model is the battery model
with its output (terminal voltage every 3 s for 3600 s: 1200 columns of one value) and stop functions; the gathered block is
[1200][1][B_total] f64.  Times are CUDA events on the launching stream, max over ranks; rank 0 prints one JSON line."""
import argparse
import ctypes
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="spm99_stop")
    ap.add_argument("--batch", type=int, default=250000)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    import diffsol_b200
    from diffsol_b200 import capi, sweeps
    from diffsol_b200 import distributed as dsbd

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    capi.require_device()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    B = args.batch
    gidx = rank + world * np.arange(B, dtype=np.int64)              # global instance ids of this shard
    current = (0.6 + 0.8 * sweeps.uniform(gidx, 0)).reshape(-1, 1)
    t_eval = np.ascontiguousarray(np.arange(1, 1201) * 3.0)
    nt = len(t_eval)
    problem = diffsol_b200.OdeBuilder().rhs_implicit(args.model).p(current).use_coloring(True).device(local_rank).build()
    solver = problem.bdf()
    nout = problem.nout
    L = capi.lib()
    vp = ctypes.c_void_p
    params_dev = torch.from_numpy(np.ascontiguousarray(current)).to(dev)
    ys_dev = torch.empty((nt * nout, B), dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream()
    capi.check(L.dsb_batch_set_params_device(solver._b, vp(params_dev.data_ptr()), B, 1, vp(stream.cuda_stream)))

    def step():
        capi.check(L.dsb_batch_solve_dense(solver._b, 0, vp(t_eval.ctypes.data), nt, vp(ys_dev.data_ptr()), vp(stream.cuda_stream)))
        if world > 1:
            return dsbd.all_gather_batch_major(ys_dev, B * world)
        return ys_dev

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    integ = []
    gathered = None
    for k in range(args.steps):
        ev[k][0].record(stream)
        gathered = step()
        ev[k][1].record(stream)
        torch.cuda.synchronize()
        integ.append(solver.last_integrator_ms())
    barrier()
    ms = [a.elapsed_time(b) for a, b in ev]
    nli = solver.sum_statistic("number_of_nonlinear_solver_iterations")
    root_idx, ncols = solver.root_info()
    status = solver.status()
    # the gathered block is in GLOBAL instance order: column b of this rank's shard is global instance rank + world * b
    # (bitwise: columns behind a root are NaN)
    ok = bool(torch.equal(gathered[:, rank::world][:, :B].contiguous().view(torch.int64), ys_dev.view(torch.int64))) if world > 1 else True
    t = torch.tensor([float(np.mean(ms)), float(np.mean(integ))], dtype=torch.float64, device=dev)
    c = torch.tensor([nli, int((status != 0).sum()), int((root_idx >= 0).sum()), int(ok)], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        cmin = c.clone()
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        dist.all_reduce(cmin, op=dist.ReduceOp.MIN)
        ok_all = bool(cmin[3].item())
    else:
        ok_all = ok
    if rank == 0:
        step_ms, integ_ms = float(t[0]), float(t[1])
        total = B * world
        v = gathered[:, 0].view(nt, nout)[:, 0].cpu().numpy()
        print(json.dumps({
            "config": "5: battery %s, batch %d per GPU x %d GPUs, BDF, 1200 output columns (terminal voltage), stop at the voltage cut-offs"
                      % (args.model, B, world),
            "n_gpus": world, "instances_total": total, "ms_per_step": step_ms, "integrator_ms": integ_ms,
            "all_gather_ms": step_ms - integ_ms, "gathered_GB": nt * nout * total * 8 / 1e9,
            "instances_per_s": total / step_ms * 1e3, "newton_iters_per_s": int(c[0]) / step_ms * 1e3,
            "failed": int(c[1]), "stopped_on_root": int(c[2]), "gather_matches_local_shard": ok_all,
            "first_instance_voltage_start_end": [float(v[0]), float(v[int(ncols[0]) - 1])]}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
