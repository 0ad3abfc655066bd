#!/usr/bin/env python
"""Kernel-only timing of every BASELINE config (C1..C5) through the product path (CUDA events around graph-replayed
launches, inputs resident).  One JSON line per case: Gpoints/s, algorithmic GB/s, fraction of the measured HBM peak.
Usage: configs.py [--generic] [--only C4,C5] [--reps N]"""
import argparse, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import deo_b200 as D
from deo_b200 import _lib

ROBIN_L, ROBIN_R = (1.0, 0.5, 0.25), (1.0, -0.5, 0.75)


def peak():
    try:
        return float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def field(shape, dtype):
    n = int(np.prod(shape))
    blk = np.random.default_rng(0).uniform(-1, 1, min(n, 1 << 26)).astype(dtype)
    host = np.empty(n, dtype=dtype)
    for i in range(0, n, blk.size):
        host[i:i + blk.size] = blk[: min(blk.size, n - i)]
    return D.DeviceArray.from_host(host.reshape(shape, order="F"))


def laplacian(shape, a, dtype, bc="robin"):
    nd = len(shape)
    h = tuple(1.0 / (s + 1) for s in shape)
    if nd == 1:
        return D.CenteredDifference(2, a, h[0], shape[0], dtype=dtype) * D.Dirichlet0BC(dtype)
    A = D.CenteredDifference[1](2, a, h[0], shape[0], dtype=dtype)
    for ax in range(2, nd + 1):
        A = A + D.CenteredDifference[ax](2, a, h[ax - 1], shape[ax - 1], dtype=dtype)
    Q = D.compose(*D.Neumann0BC(dtype, h, 1, shape)) if bc == "neumann" else D.compose(*D.RobinBC(ROBIN_L, ROBIN_R, h, 1, shape, dtype=dtype))
    return A * Q


def c4_families(shape, dtype):
    """BASELINE config 4: non-uniform grid, centered (2,4) + centered (1,4) + upwind (1,2) with a mixed-sign
    coefficient vector per axis, RobinBC from the same spacing vectors."""
    hs = [1.0 / (s + 1) for s in shape]
    dxs = [(h * (1 + 0.3 * np.sin(2 * np.pi * np.arange(1, s + 2) / (s + 1)))).astype(dtype) for s, h in zip(shape, hs)]
    cs = [np.sin(6 * np.pi * np.arange(1, s + 1) / s).astype(dtype) for s in shape]
    axes = range(1, len(shape) + 1)
    fam = {
        "lap": [D.CenteredDifference[ax](2, 4, dxs[ax - 1], shape[ax - 1], dtype=dtype) for ax in axes],
        "grad": [D.CenteredDifference[ax](1, 4, dxs[ax - 1], shape[ax - 1], dtype=dtype) for ax in axes],
        "adv": [D.UpwindDifference[ax](1, 2, dxs[ax - 1], shape[ax - 1], cs[ax - 1], dtype=dtype) for ax in axes],
    }
    fam["all9"] = fam["lap"] + fam["grad"] + fam["adv"]
    Q = D.compose(*D.RobinBC(ROBIN_L, ROBIN_R, dxs, 1, shape, dtype=dtype))
    out = {}
    for name, ops in fam.items():
        A = ops[0]
        for o in ops[1:]:
            A = A + o
        out[name] = A * Q
    return out


def table_bytes(G, dtype):
    """Coefficient traffic counted once per application: per-row weight tables and coefficient vectors."""
    from deo_b200.apply import _terms
    es = np.dtype(dtype).itemsize
    b = 0
    for L, _ in _terms(G):
        if L.nonuniform:
            b += (np.asarray(L.stencil_coefs).size + np.asarray(L.low_boundary_coefs).size + np.asarray(L.high_boundary_coefs).size) * es
        if np.ptp(np.asarray(L.coefficients)) != 0:
            b += L.len * es
    return b


def run(name, G, shape, dtype, flags, reps, u=None):
    dtype = np.dtype(dtype)
    plan = D.build_plans(G, shape, shape, dtype, flags=flags)[0][0]
    u = u if u is not None else field(shape, dtype)
    du = D.DeviceArray(shape, dtype)
    for _ in range(3):
        plan.apply(du, u)
    D.sync()
    ms = min(plan.time(du, u, reps) for _ in range(3))
    n = float(np.prod(shape))
    alg = 2 * dtype.itemsize * n + table_bytes(G, dtype)
    gbs = alg / (ms * 1e-3) / 1e9
    rec = {"case": name, "shape": list(shape), "dtype": dtype.name, "kernel": plan.info[0], "ms": ms, "gpoints_per_s": n / ms / 1e6,
           "algorithmic_GBps": gbs, "frac_measured_peak": gbs / peak(), "frac_8TBs": gbs / 8000.0, "algorithmic_bytes": alg}
    print(json.dumps(rec), flush=True)
    return u


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--generic", action="store_true")
    ap.add_argument("--only", default="")
    ap.add_argument("--reps", type=int, default=20)
    args = ap.parse_args()
    flags = _lib.DEO_FLAG_FORCE_GENERIC if args.generic else 0
    only = set(filter(None, args.only.split(",")))
    want = lambda c: not only or c in only
    if want("C1"):
        run("C1 1-D (2,2) Dirichlet0 N=1e6", laplacian((10 ** 6,), 2, np.float64), (10 ** 6,), np.float64, flags, 1000)
    if want("C2"):
        run("C2 2-D (2,4) Robin 8192^2", laplacian((8192, 8192), 4, np.float64), (8192, 8192), np.float64, flags, args.reps)
    if want("C3"):
        for dt in (np.float64, np.float32):
            run("C3 3-D (2,6) Neumann0 512^3", laplacian((512,) * 3, 6, dt, "neumann"), (512,) * 3, dt, flags, args.reps)
    if want("C4"):
        u = None
        for fam, G in c4_families((512,) * 3, np.float64).items():
            u = run(f"C4 non-uniform 512^3 {fam}", G, (512,) * 3, np.float64, flags, args.reps, u)
        del u
    if want("C5"):
        run("C5 3-D (2,4) Robin 1024^3", laplacian((1024,) * 3, 4, np.float64), (1024,) * 3, np.float64, flags, args.reps)
        run("C5 3-D (2,2) Robin 1024^3", laplacian((1024,) * 3, 2, np.float64), (1024,) * 3, np.float64, flags, args.reps)
