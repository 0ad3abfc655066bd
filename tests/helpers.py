"""Shared helpers for the parity tests: build the same operator on the product side (deo_b200 host
mirror -> CUDA library) and on the oracle side, and compare."""
import numpy as np

TOL = {np.dtype(np.float64): 1e-13, np.dtype(np.float32): 1e-5}   # north_star tolerances, max over all points


def rel_err(got, want):
    """max|gpu - oracle| / max|oracle| over ALL points including boundary rows."""
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, (got.shape, want.shape)
    denom = np.abs(want).max()
    if denom == 0:
        denom = 1.0
    return float(np.abs(got - want).max() / denom)


def assert_close(got, want, dtype, what=""):
    e = rel_err(got, want)
    assert e <= TOL[np.dtype(dtype)], f"{what}: rel err {e:.3e} > {TOL[np.dtype(dtype)]:.0e}"
    return e


def uniform_field(shape, dtype, seed=0):
    """Gate inputs: iid Uniform(-1,1), generated in Float64 then cast."""
    rng = np.random.default_rng(seed)
    return np.asfortranarray(rng.uniform(-1, 1, shape).astype(dtype))


def nonuniform_dx(n, h, dtype):
    """dx_a[i] = h*(1 + 0.3 sin(2 pi i / (n+1))), n+1 spacings (BASELINE config C4)."""
    i = np.arange(1, n + 2)
    return (h * (1 + 0.3 * np.sin(2 * np.pi * i / (n + 1)))).astype(dtype)


def make_pair(kind, d, a, dx, n, coeff=1, axis=1, offside=0, dtype=np.float64):
    """-> (deo_b200 operator, oracle operator) built from identical arguments."""
    import deo_b200 as D
    from oracle import oracle as O
    if kind == "centered":
        return (D.CenteredDifference[axis](d, a, dx, n, coeff, dtype=dtype),
                O.CenteredDifference(d, a, dx, n, coeff, axis=axis, dtype=dtype))
    return (D.UpwindDifference[axis](d, a, dx, n, coeff, offside=offside, dtype=dtype),
            O.UpwindDifference(d, a, dx, n, coeff, offside=offside, axis=axis, dtype=dtype))


def bc_pair(spec, dx, dtype):
    """spec: ('dirichlet0',) | ('dirichlet', al, ar) | ('neumann', (al, ar), order) | ('robin', l, r, order)
    | ('general', alpha_l, alpha_r, order) | ('periodic',)"""
    import deo_b200 as D
    from oracle import oracle as O
    k = spec[0]
    if k == "dirichlet0":
        return D.Dirichlet0BC(dtype), O.Dirichlet0BC(dtype)
    if k == "dirichlet":
        return D.DirichletBC(spec[1], spec[2], dtype=dtype), O.DirichletBC(spec[1], spec[2], dtype)
    if k == "neumann":
        return D.NeumannBC(spec[1], dx, spec[2], dtype=dtype), O.NeumannBC(spec[1], dx, spec[2], dtype)
    if k == "robin":
        return D.RobinBC(spec[1], spec[2], dx, spec[3], dtype=dtype), O.RobinBC(spec[1], spec[2], dx, spec[3], dtype)
    if k == "general":
        return D.GeneralBC(spec[1], spec[2], dx, spec[3], dtype=dtype), O.GeneralBC(spec[1], spec[2], dx, spec[3], dtype)
    if k == "periodic":
        return D.PeriodicBC(dtype), O.PeriodicBC(dtype)
    raise ValueError(k)
