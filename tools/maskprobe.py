#!/usr/bin/env python
"""Time the tiled kernel per operator mask (which axes carry an operator) on one shape: the memory traffic is the same for
every mask, so the differences are the cost of each axis' arithmetic / shared-memory loads.
Usage: maskprobe.py SHAPE APPROX DTYPE [masks, default 1,2,4,3,5,6,7]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import deo_b200 as D

shape = tuple(int(v) for v in sys.argv[1].split("x"))
a = int(sys.argv[2]); dtype = np.dtype(sys.argv[3])
masks = [int(m) for m in (sys.argv[4] if len(sys.argv) > 4 else "1,2,4,3,5,6,7").split(",")]
h = tuple(1.0 / (s + 1) for s in shape)
n = int(np.prod(shape))
host = np.resize(np.random.default_rng(0).uniform(-1, 1, 1 << 22).astype(dtype), n)
u = D.DeviceArray.from_host(host.reshape(shape, order="F")); du = D.DeviceArray(shape, dtype)
for m in masks:
    axes = [ax for ax in range(1, len(shape) + 1) if (m >> (ax - 1)) & 1]
    A = None
    for ax in axes:
        op = D.CenteredDifference[ax](2, a, h[ax - 1], shape[ax - 1], dtype=dtype)
        A = op if A is None else A + op
    Q = D.compose(*D.RobinBC((1.0, 0.5, 0.25), (1.0, -0.5, 0.75), h, 1, shape, dtype=dtype))
    plan = D.build_plans(A * Q, shape, shape, dtype)[0][0]
    for _ in range(3): plan.apply(du, u)
    D.sync()
    ms = min(plan.time(du, u, 10) for _ in range(3))
    print(f"{sys.argv[1]} {dtype.name} a={a} mask={m} {plan.info[0]} {ms:8.4f} ms {n / ms / 1e6:7.1f} Gpts/s ({2 * dtype.itemsize * n / ms / 1e6 / 6559.4:.3f})", flush=True)
