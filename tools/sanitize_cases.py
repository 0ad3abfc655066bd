#!/usr/bin/env python
"""Small parity shapes for compute-sanitizer (memcheck / racecheck / synccheck): tiled CONST and TABLE plans in 3-D, a 2-D
strip, accumulate + fused AXPY, per-pencil BC tables, a pre-padded input; every result is compared with the per-point kernel."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import deo_b200 as D

ROBIN = ((1.0, 0.5, 0.25), (1.0, -0.5, 0.75))
GEN = D._lib.DEO_FLAG_FORCE_GENERIC


def lap(shape, a, dtype, dxs=None):
    h = tuple(1.0 / (s + 1) for s in shape) if dxs is None else dxs
    A = D.CenteredDifference[1](2, a, h[0], shape[0], dtype=dtype)
    for ax in range(2, len(shape) + 1):
        A = A + D.CenteredDifference[ax](2, a, h[ax - 1], shape[ax - 1], dtype=dtype)
    return A, h


def check(name, G, u, shape, dtype):
    ud = D.DeviceArray.from_host(u)
    a = D.mul_alloc(G, ud).to_host()
    b = D.mul_alloc(G, ud, flags=GEN).to_host()
    kern = D.apply._get_plans(G, shape, shape, dtype, False, 0)[0][0].info[0]
    err = np.abs(a.astype(np.float64) - b).max() / np.abs(b).max()
    print(f"{name:40s} kernel={kern:10s} rel diff vs per-point kernel {err:.2e}", flush=True)
    assert err < (1e-13 if dtype == np.float64 else 1e-5)


rng = np.random.default_rng(0)
for dtype in (np.float64, np.float32):
    for shape, a in (((136, 70, 40), 4), ((72, 44, 30), 6), ((264, 60), 4)):
        u = np.asfortranarray(rng.uniform(-1, 1, shape).astype(dtype))
        A, h = lap(shape, a, dtype)
        G = A * D.compose(*D.RobinBC(*ROBIN, h, 1, shape, dtype=dtype))
        check(f"CONST {shape} a={a} {np.dtype(dtype).name}", G, u, shape, dtype)
        dxs = [(hh * (1 + 0.3 * np.sin(2 * np.pi * np.arange(1, s + 2) / (s + 1)))).astype(dtype) for s, hh in zip(shape, h)]
        c = np.sin(6 * np.pi * np.arange(1, shape[-1] + 1) / shape[-1]).astype(dtype)
        At, _ = lap(shape, 4, dtype, dxs)
        At = At + D.UpwindDifference[len(shape)](1, 2, dxs[-1], shape[-1], c, dtype=dtype)
        Gt = At * D.compose(*D.RobinBC(*ROBIN, dxs, 1, shape, dtype=dtype))
        check(f"TABLE {shape} {np.dtype(dtype).name}", Gt, u, shape, dtype)
        # accumulate + fused AXPY
        ud = D.DeviceArray.from_host(u)
        out = D.DeviceArray(shape, dtype); D.step_(out, G, ud, 1e-6)
        acc = D.DeviceArray.from_host(u); D.mul_(acc, G, ud, overwrite=False)
        D.sync()
print("OK")
