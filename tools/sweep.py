#!/usr/bin/env python
"""Knob sweep of the tiled kernel on one workload (kernel-only timing with CUDA events, or one launch per
configuration when run under ncu with SWEEP_NCU=1).  Usage: sweep.py SHAPE APPROX DTYPE 'K=V,K=V;K=V...'"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import deo_b200 as D

shape = tuple(int(v) for v in sys.argv[1].split("x"))
a = int(sys.argv[2])
dtype = np.dtype(sys.argv[3])
configs = [dict(kv.split("=") for kv in c.split(",") if kv) for c in sys.argv[4].split(";")]
under_ncu = os.environ.get("SWEEP_NCU") == "1"
h = tuple(1.0 / (s + 1) for s in shape)
n = int(np.prod(shape))
blk = np.random.default_rng(0).uniform(-1, 1, min(n, 1 << 26)).astype(dtype)
host = np.empty(n, dtype=dtype)
for i in range(0, n, blk.size):
    host[i:i + blk.size] = blk[: min(blk.size, n - i)]
u = D.DeviceArray.from_host(host.reshape(shape, order="F"))
du = D.DeviceArray(shape, dtype)
del host, blk
for cfg in configs:
    for k in ("DEO_TMA_L2PROMO", "DEO_STAR_ZCHUNK", "DEO_STAR_PY", "DEO_STAR_NWY", "DEO_STAR_V"):
        os.environ.pop(k, None)
    os.environ.update(cfg)
    A = D.CenteredDifference[1](2, a, h[0], shape[0], dtype=dtype)
    for ax in range(2, len(shape) + 1):
        A = A + D.CenteredDifference[ax](2, a, h[ax - 1], shape[ax - 1], dtype=dtype)
    Q = D.compose(*D.RobinBC((1.0, 0.5, 0.25), (1.0, -0.5, 0.75), h, 1, shape, dtype=dtype))
    plan = D.build_plans(A * Q, shape, shape, dtype)[0][0]
    if under_ncu:
        plan.apply(du, u); D.sync()
        print("NCU", cfg, flush=True)
        continue
    for _ in range(3):
        plan.apply(du, u)
    D.sync()
    ms = min(plan.time(du, u, 10) for _ in range(3))
    gbs = 2 * dtype.itemsize * n / (ms * 1e-3) / 1e9
    print(f"{sys.argv[1]} {dtype.name} a={a} {plan.info[0]} {cfg} {ms:8.4f} ms {n / ms / 1e6:7.1f} Gpts/s {gbs:7.0f} GB/s ({gbs / 6559.4:.3f})", flush=True)
