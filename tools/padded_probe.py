#!/usr/bin/env python
"""Kernel-only timing of mul!(x_temp, Dxx+Dyy[+Dzz], M) on a dense PRE-PADDED array M (derivative_operator_functions.jl:203,:466)
and of one operator along the contiguous axis of it (the shape every Gradient / Divergence / Curl component has).
Usage: padded_probe.py SHAPE APPROX DTYPE   (SHAPE = extents of the output)"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import deo_b200 as D

shape = tuple(int(v) for v in sys.argv[1].split("x"))
a = int(sys.argv[2]); dtype = np.dtype(sys.argv[3])
h = tuple(1.0 / (s + 1) for s in shape)
pshape = tuple(s + 2 for s in shape)
n = int(np.prod(shape))
host = np.resize(np.random.default_rng(0).uniform(-1, 1, 1 << 22).astype(dtype), int(np.prod(pshape)))
M = D.DeviceArray.from_host(host.reshape(pshape, order="F")); du = D.DeviceArray(shape, dtype)
ops = [D.CenteredDifference[ax](2, a, h[ax - 1], shape[ax - 1], dtype=dtype) for ax in range(1, len(shape) + 1)]
for name, A in (("sum of all axes", sum(ops[1:], ops[0])), ("x axis only", ops[0])):
    plan = D.build_plans(A, shape, pshape, dtype)[0][0]
    for _ in range(3): plan.apply(du, M)
    D.sync()
    ms = min(plan.time(du, M, 10) for _ in range(3))
    print(f"pre-padded {sys.argv[1]} {dtype.name} a={a} {name:16s} {plan.info[0]} {ms:8.4f} ms {n / ms / 1e6:7.1f} Gpts/s ({2 * dtype.itemsize * n / ms / 1e6 / 6559.4:.3f})", flush=True)
