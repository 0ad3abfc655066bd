#!/usr/bin/env python
"""Probe: cache-resident vs DRAM-resident problem sizes for the fused Laplacian (kernel-only timing)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import deo_b200 as D

def run(shape, axes, a, dtype=np.float64, reps=20):
    h = tuple(1.0 / (s + 1) for s in shape)
    ops = [D.CenteredDifference[ax](2, a, h[ax - 1], shape[ax - 1], dtype=dtype) for ax in axes]
    A = ops[0]
    for o in ops[1:]:
        A = A + o
    Q = D.compose(*D.RobinBC((1.0, 0.5, 0.25), (1.0, -0.5, 0.75), h, 1, shape, dtype=dtype))
    plan = D.build_plans(A * Q, shape, shape, dtype)[0][0]
    u = D.DeviceArray.from_host(np.asfortranarray(np.random.default_rng(0).uniform(-1, 1, shape).astype(dtype)))
    du = D.DeviceArray(shape, dtype)
    for _ in range(3):
        plan.apply(du, u)
    D.sync()
    ms = plan.time(du, u, reps)
    pts = float(np.prod(shape))
    gbs = 2 * np.dtype(dtype).itemsize * pts / (ms * 1e-3) / 1e9
    print(f"{'x'.join(map(str, shape)):>16} {np.dtype(dtype).name} axes={axes} a={a} {plan.info[0]:8s} {ms:8.4f} ms {pts / ms / 1e6:7.1f} Gpts/s {gbs:7.0f} GB/s ({gbs / 6559.4:.2f})", flush=True)

if __name__ == "__main__":
    for shape in [(1024, 1024, 128), (512, 512, 128), (512, 256, 64), (1024, 592, 64)]:
        for axes, a in [((3,), 4), ((1, 2, 3), 2), ((1, 2, 3), 4), ((1, 2, 3), 6)]:
            run(shape, axes, a)
