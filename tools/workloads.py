"""The BASELINE.json configurations as operator builders, shared by bench.py, the full-size parity tests and
the multi-GPU checker.  Nothing here imports the product or the oracle: the caller passes the module (`D` = deo_b200 host
mirror, `O` = oracle front-end), so the product path never sees the oracle.

Fields are reproducible plane by plane (`field_planes`): plane k of the last axis only depends on (seed, k), so every rank
of a slab-decomposed run -- and the CPU oracle checking it -- can regenerate exactly the planes it needs.
"""
import numpy as np

WORKLOADS = {
    # name: (shape, approximation order, dtype, description)
    "C5": ((1024, 1024, 1024), 4, np.float64, "C5: 3-D Laplacian Dxx+Dyy+Dzz CenteredDifference(2,4) + RobinBC, 1024^3 Float64, slabs along dim 3"),
    "C3": ((512, 512, 512), 6, np.float64, "C3: 3-D Laplacian (2,6) + Neumann MultiDimBC, 512^3 Float64"),
    "C3f32": ((512, 512, 512), 6, np.float32, "C3: 3-D Laplacian (2,6) + Neumann MultiDimBC, 512^3 Float32"),
    "C2": ((8192, 8192), 4, np.float64, "C2: 2-D Laplacian Dxx+Dyy (2,4) + RobinBC, 8192^2 Float64"),
    "C1": ((10 ** 6,), 2, np.float64, "C1: 1-D heat-equation Laplacian CenteredDifference(2,2)*Dirichlet0BC, N=1e6 Float64"),
    "C4": ((512, 512, 512), 4, np.float64, "C4: non-uniform 512^3 Float64, sum over axes of CenteredDifference(2,4) + CenteredDifference(1,4) + "
           "UpwindDifference(1,2) with a mixed-sign coefficient vector, RobinBC from the same spacings (9 operators, one pass)"),
}
ROBIN_L, ROBIN_R = (1.0, 0.5, 0.25), (1.0, -0.5, 0.75)


def c4_inputs(shape, dtype):
    """Spacing vectors (len+1 entries) and coefficient vectors of BASELINE config 4."""
    hs = [1.0 / (s + 1) for s in shape]
    dxs = [(h * (1 + 0.3 * np.sin(2 * np.pi * np.arange(1, s + 2) / (s + 1)))).astype(dtype) for s, h in zip(shape, hs)]
    cs = [np.sin(6 * np.pi * np.arange(1, s + 1) / s).astype(dtype) for s in shape]
    return dxs, cs


def build_operator(D, name, shape, dtype):
    """The product-side operator A*Q of a workload (host mirror of the reference constructors)."""
    a = WORKLOADS[name][1]
    nd = len(shape)
    h = tuple(1.0 / (s + 1) for s in shape)
    if name == "C4":
        dxs, cs = c4_inputs(shape, dtype)
        ops = [D.CenteredDifference[ax](2, 4, dxs[ax - 1], shape[ax - 1], dtype=dtype) for ax in range(1, nd + 1)] + \
              [D.CenteredDifference[ax](1, 4, dxs[ax - 1], shape[ax - 1], dtype=dtype) for ax in range(1, nd + 1)] + \
              [D.UpwindDifference[ax](1, 2, dxs[ax - 1], shape[ax - 1], cs[ax - 1], dtype=dtype) for ax in range(1, nd + 1)]
        A = ops[0]
        for o in ops[1:]:
            A = A + o
        return A * D.compose(*D.RobinBC(ROBIN_L, ROBIN_R, dxs, 1, shape, dtype=dtype))
    if nd == 1:
        return D.CenteredDifference(2, a, h[0], shape[0], dtype=dtype) * D.Dirichlet0BC(dtype)
    A = D.CenteredDifference[1](2, a, h[0], shape[0], dtype=dtype)
    for ax in range(2, nd + 1):
        A = A + D.CenteredDifference[ax](2, a, h[ax - 1], shape[ax - 1], dtype=dtype)
    if name.startswith("C3"):
        Q = D.compose(*D.Neumann0BC(dtype, h, 1, shape))
    else:
        Q = D.compose(*D.RobinBC(ROBIN_L, ROBIN_R, h, 1, shape, dtype=dtype))
    return A * Q


def build_oracle(O, name, gshape, dtype, last=None):
    """Oracle-side operators and BCs of a workload on the global grid `gshape`.  `last=(a, b)` builds the SUB-PROBLEM on
    rows [a, b) of the last axis: same spacing h (taken from the global extent), spacing / coefficient vectors sliced, the
    last operator's len = b - a.  Rows of the sub-problem further than the stencils' reach from a cut equal the global
    problem's rows (interior stencils only see local data); rows at a real face (a == 0 or b == n) are exact too."""
    a_ord = WORKLOADS[name][1]
    nd = len(gshape)
    h = tuple(1.0 / (s + 1) for s in gshape)
    lo, hi = (0, gshape[-1]) if last is None else last
    shape = tuple(gshape[:-1]) + (hi - lo,)
    if name == "C4":
        dxs, cs = c4_inputs(gshape, dtype)
        dxs[-1] = dxs[-1][lo:hi + 1]          # spacings between the padded nodes lo .. hi+1
        cs[-1] = cs[-1][lo:hi]
        ops = [O.CenteredDifference(2, 4, dxs[ax], shape[ax], axis=ax + 1, dtype=dtype) for ax in range(nd)] + \
              [O.CenteredDifference(1, 4, dxs[ax], shape[ax], axis=ax + 1, dtype=dtype) for ax in range(nd)] + \
              [O.UpwindDifference(1, 2, dxs[ax], shape[ax], cs[ax], axis=ax + 1, dtype=dtype) for ax in range(nd)]
        return ops, {ax + 1: O.RobinBC(ROBIN_L, ROBIN_R, dxs[ax], 1, dtype) for ax in range(nd)}
    ops = [O.CenteredDifference(2, a_ord, h[ax], shape[ax], axis=ax + 1, dtype=dtype) for ax in range(nd)]
    if nd == 1:
        bcs = {1: O.Dirichlet0BC(dtype)}
    elif name.startswith("C3"):
        bcs = {ax + 1: O.Neumann0BC(h[ax], 1, dtype) for ax in range(nd)}
    else:
        bcs = {ax + 1: O.RobinBC(ROBIN_L, ROBIN_R, h[ax], 1, dtype) for ax in range(nd)}
    return ops, bcs


ORACLE_MARGIN = 20    # rows of a sub-problem next to a cut that are discarded (> 2 * boundary_stencil_length of every config)


def field_planes(gshape, dtype, z0, z1, seed=1234):
    """Planes [z0, z1) of the last axis of the global synthetic field: iid Uniform(-1,1) generated in Float64 then cast
    (SURVEY 8d), plane k from the generator seeded (seed, k).  1-D fields are one 'plane' (k = 0)."""
    if len(gshape) == 1:
        return np.random.default_rng([seed, 0]).uniform(-1, 1, gshape).astype(dtype)[z0:z1]
    out = np.empty(tuple(gshape[:-1]) + (z1 - z0,), dtype=dtype, order="F")
    for k in range(z0, z1):
        out[..., k - z0] = np.random.default_rng([seed, k]).uniform(-1, 1, gshape[:-1][::-1]).T
    return out


def oracle_rows(O, name, gshape, dtype, z0, z1, u_of, nthreads=1):
    """The oracle's result on rows [z0, z1) of the last axis of the global problem, from a sub-problem with ORACLE_MARGIN
    extra rows on each side (clipped at the real faces).  `u_of(a, b)` returns the global field's planes [a, b)."""
    n = gshape[-1]
    if len(gshape) == 1:
        ops, bcs = build_oracle(O, name, gshape, dtype)
        return O.apply_sum(ops, u_of(0, n), bcs, nthreads=nthreads)[z0:z1]
    a, b = max(z0 - ORACLE_MARGIN, 0), min(z1 + ORACLE_MARGIN, n)
    ops, bcs = build_oracle(O, name, gshape, dtype, last=(a, b))
    want = O.apply_sum(ops, u_of(a, b), bcs, nthreads=nthreads)
    return want[..., z0 - a:z1 - a]


def check_rows(O, name, gshape, dtype, got_rows, z0, z1, u_of, nthreads=1):
    """-> (max|got - oracle|, max|oracle|, boundary-only max abs error) over rows [z0, z1) of the last axis;
    `got_rows` holds exactly those rows.  'Boundary' = points within 4 rows of any face of the global grid."""
    want = oracle_rows(O, name, gshape, dtype, z0, z1, u_of, nthreads).astype(np.float64)
    diff = np.abs(np.asarray(got_rows, dtype=np.float64) - want)
    nd = len(gshape)
    bmask = np.zeros(diff.shape, dtype=bool)
    for ax in range(nd):
        idx = np.arange(diff.shape[ax]) + (z0 if ax == nd - 1 else 0)
        near = (idx < 4) | (idx >= gshape[ax] - 4)
        sh = [1] * nd
        sh[ax] = -1
        bmask |= near.reshape(sh)
    berr = float(diff[bmask].max()) if bmask.any() else 0.0
    return float(diff.max()), float(np.abs(want).max()), berr
