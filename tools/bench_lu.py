#!/usr/bin/env python
"""Micro-benchmark of the block-cooperative batched LU pair (instance-major storage): factor and solve
throughput and HBM-roofline fraction (algorithmic bytes: factor 16 n^2 + 4 n, solve 8 n^2 + 20 n per instance)."""
import ctypes
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffsol_b200 import capi  # noqa: E402

capi.require_device()
L = capi.lib()
vp = ctypes.c_void_p
peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
dev = torch.device("cuda:0")
for n, B in ((32, 65536), (64, 32768), (128, 16384), (256, 16384)):
    g = torch.Generator(device=dev).manual_seed(n)
    a0 = torch.randn((B, n, n), dtype=torch.float64, device=dev, generator=g)
    a0 += torch.eye(n, dtype=torch.float64, device=dev) * 4.0
    a = a0.clone()
    rhs = torch.randn((B, n), dtype=torch.float64, device=dev, generator=g)
    piv = torch.zeros((B, n), dtype=torch.int32, device=dev)
    info = torch.zeros(B, dtype=torch.int32, device=dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    tf, ts = [], []
    for it in range(4):
        a.copy_(a0)
        torch.cuda.synchronize()
        ev[0].record()
        capi.check(L.dsb_lu_factor_instance_major(vp(a.data_ptr()), n, B, vp(piv.data_ptr()), vp(info.data_ptr()), None))
        ev[1].record()
        capi.check(L.dsb_lu_solve_instance_major(vp(a.data_ptr()), vp(piv.data_ptr()), vp(rhs.data_ptr()), n, B, vp(info.data_ptr()), None))
        ev[2].record()
        torch.cuda.synchronize()
        if it:
            tf.append(ev[0].elapsed_time(ev[1])); ts.append(ev[1].elapsed_time(ev[2]))
    tf, ts = min(tf), min(ts)
    bf, bs = B * (16 * n * n + 4 * n), B * (8 * n * n + 20 * n)
    print(json.dumps({"n": n, "batch": B, "factor_ms": tf, "solve_ms": ts,
                      "factor_per_s": B / tf * 1e3, "solve_per_s": B / ts * 1e3,
                      "factor_GBps": bf / tf / 1e6, "solve_GBps": bs / ts / 1e6,
                      "factor_frac_hbm": bf / tf / 1e6 / peak, "solve_frac_hbm": bs / ts / 1e6 / peak,
                      "factor_GFLOPs": B * (2 * n ** 3 / 3) / tf / 1e6}))
