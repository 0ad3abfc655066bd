"""Pins the CPU oracle (oracle/) against the golden vectors held by the
reference's own tests (tests/golden/reference_kats.json, transcribed by
tests/golden/make_golden.py with file:line citations).  CPU only."""
import numpy as np
import pytest

from oracle import oracle as O

RTOL = 1e-8   # the reference tests use `≈` (rtol = sqrt(eps) ~ 1.5e-8); we hold the oracle tighter


def _op(e, dtype=np.float64):
    dx = e["dx"] if np.ndim(e["dx"]) == 0 else np.asarray(e["dx"])
    if e["kind"] == "centered":
        return O.CenteredDifference(e["d"], e["a"], dx, e["n"], e["coeff"], dtype=dtype)
    return O.UpwindDifference(e["d"], e["a"], dx, e["n"], e["coeff"], offside=e.get("offside", 0), dtype=dtype)


def test_operator_matrices(golden):
    assert len(golden["operator_matrices"]) >= 30
    for e in golden["operator_matrices"]:
        L = _op(e)
        got = L.to_matrix()
        want = np.asarray(e["matrix"])
        if "rows" in e:
            got = got[e["rows"][0]:e["rows"][1]]
        scale = np.abs(want).max()
        assert got.shape == want.shape, e["cite"]
        assert np.abs(got - want).max() <= RTOL * scale, e["cite"]


def test_bpv_equals_plain_on_fixtures(golden):
    """convolutions.jl(test):45-81: DerivativeOperator*BoundaryPaddedVector == stencil*[l;u;r]."""
    rng = np.random.default_rng(1)
    for e in golden["operator_matrices"]:
        L = _op(e)
        n = e["n"]
        l, r = rng.random(2)
        u = rng.random(n)
        want = L.to_matrix() @ np.concatenate([[l], u, [r]])
        got = L.mul1d_bpv(l, r, u)
        assert np.abs(got - want).max() <= 1e-12 * max(1.0, np.abs(want).max()), e["cite"]


def test_interior_weights(golden):
    g = golden["interior_weights"]
    for e in g["centered"]:
        L = O.CenteredDifference(e["d"], e["a"], 1.0, 12)
        assert np.allclose(L.stencil_coefs, e["w"], atol=1e-10, rtol=0)
    for e in g["upwind"]:
        L = O.UpwindDifference(e["d"], e["a"], 1.0, 12)
        assert np.allclose(L.stencil_coefs, e["w"], atol=1e-10, rtol=0)


def test_vector_coefficients(golden):
    g = golden["vector_coefficients"]
    x = np.asarray(g["x"]); c = np.asarray(g["c"])
    A = O.CenteredDifference(2, 2, 1.0, 3)
    B = O.UpwindDifference(1, 1, 1.0, 3, 1.0)
    Cc = O.UpwindDifference(1, 1, 1.0, 3, [2.0, 3.0, 4.0])
    for mk in (lambda: O.CenteredDifference(2, 2, 1.0, 3), lambda: O.UpwindDifference(1, 1, 1.0, 3, 1.0),
               lambda: O.UpwindDifference(1, 1, 1.0, 3, [2.0, 3.0, 4.0])):
        base = mk(); scaled = mk().scale(c)
        assert np.allclose(base.mul1d_plain(x) * c, scaled.mul1d_plain(x), rtol=1e-12)
    assert np.array_equal(A.scale(c).coefficients, c)
    assert np.array_equal(B.scale(c).coefficients, c)
    assert np.array_equal(Cc.scale(c).coefficients, g["cC_coefficients"])


def test_robin_order1_closed_forms(golden):
    for case in golden["robin_order1"]["cases"]:
        al, bl, cl, dx, ar, br, cr = (case[k] for k in ("al", "bl", "cl", "dx", "ar", "br", "cr"))
        u = np.asarray(case["u"])
        want_l = (cl - (bl / dx) * u[0]) / (al - bl / dx)
        want_r = (cr + (br / dx) * u[-1]) / (ar + br / dx)
        for Q in (O.RobinBC((al, bl, cl), (ar, br, cr), dx),
                  O.RobinBC((al, bl, cl), (ar, br, cr), dx * np.ones(len(u)))):
            l, r = Q.ghosts(u)
            assert np.isclose(l, want_l, rtol=1e-10) and np.isclose(r, want_r, rtol=1e-10)
            # Q_L first/last row and Q_b (robin.jl:21-26)
            assert np.isclose(Q.a_l[0], 1 / (1 - al * dx / bl), rtol=1e-10)
            assert np.isclose(Q.a_r[0], 1 / (1 + ar * dx / br), rtol=1e-10)
            assert np.isclose(Q.b_l[0], cl / (al - bl / dx), rtol=1e-10)
            assert np.isclose(Q.b_r[0], cr / (ar + br / dx), rtol=1e-10)


def test_robin_order3_and_general(golden):
    g = golden["robin_order3"]
    u = np.asarray(g["u"])
    Q = O.RobinBC(g["l"], g["r"], g["dx"], g["order"])
    l, r = Q.ghosts(u)
    assert np.isclose(l, g["u0"], rtol=1e-12) and np.isclose(r, g["uend"], rtol=1e-12)
    for dx in (1.0, np.ones(10)):
        G = O.GeneralBC(g["general_alpha"], g["general_alpha"], dx, 3)
        l, r = G.ghosts(u)
        assert np.isclose(l, g["u0"], rtol=1e-12) and np.isclose(r, g["uend"], rtol=1e-12)


def test_dirichlet_neumann_special_cases():
    # bc_operators.jl:173-183; Dirichlet relies on 1/0 = Inf giving a = (-0.0, 0.0), b = gamma
    D = O.DirichletBC(2.5, -1.5)
    assert np.all(D.a_l == 0) and np.all(D.a_r == 0) and D.b_l[0] == 2.5 and D.b_r[0] == -1.5
    N = O.NeumannBC((0.3, -0.2), 0.1, 1)
    assert N.a_l[0] == 1 and N.a_r[0] == 1
    assert np.isclose(N.b_l[0], -0.3 * 0.1) and np.isclose(N.b_r[0], -0.2 * 0.1)


def _bc_from(spec, dx):
    if spec["type"] == "neumann0":
        return O.Neumann0BC(dx, spec["order"])
    return O.RobinBC(spec["l"], spec["r"], dx, spec["order"])


def test_ghost_operator_matrices(golden):
    """Array(L*Q)[1] fixtures of BasicSDOExamples.jl: linear part of u -> L*(Q*u), recovered by probing."""
    g = golden["ghost_operator_matrices"]
    dx, M = g["dx"], g["M"]
    for case in g["cases"]:
        Q = _bc_from(case["bc"], dx)
        total = np.zeros((M, M))
        for t in case["terms"]:
            mk = O.CenteredDifference if t["kind"] == "centered" else O.UpwindDifference
            L = mk(t["d"], t["a"], dx, M, t["coeff"])
            if "scale" in t:
                L.scale(t["scale"])
            affine = O.apply_axis(L, np.zeros(M), Q)
            lin = np.stack([O.apply_axis(L, e, Q) - affine for e in np.eye(M)], axis=1)
            total += lin
        total *= case.get("scale", 1.0)
        want = np.asarray(case["matrix"])
        assert np.abs(total - want).max() <= 1e-9 * np.abs(want).max() + 1e-12, case["name"]


def test_nd_axis(golden):
    from tests.golden.make_golden import fourth_deriv_approx_stencil, second_derivative_stencil
    for case in golden["nd_axis"]["cases"]:
        shape, ax, n = case["shape"], case["axis"], case["n"]
        idx = np.arange(1, shape[ax - 1] + 1) * 0.1
        line = np.sin(idx) if case["field"] == "sin" else np.cos(idx)
        bshape = [1] * len(shape); bshape[ax - 1] = -1
        M = np.broadcast_to(line.reshape(bshape), shape).copy(order="F")
        L = O.CenteredDifference(case["d"], case["a"], case["dx"], n, axis=ax)
        got = O.apply_axis(L, M)
        S = fourth_deriv_approx_stencil(n) if case["d"] == 4 else second_derivative_stencil(n)
        correct = (1.0 / case["dx"] ** case["d"]) * (S @ line)
        want = np.broadcast_to(correct.reshape(bshape), got.shape)
        assert got.shape == tuple(s - 2 if i == ax - 1 else s for i, s in enumerate(shape))
        assert np.allclose(got, want, rtol=1e-7, atol=1e-7 * np.abs(correct).max())


def test_fornberg_f32_is_f32_arithmetic():
    w32 = O.calculate_weights(2, 0.0, np.arange(-3, 4), dtype=np.float32)
    w64 = O.calculate_weights(2, 0.0, np.arange(-3, 4), dtype=np.float64)
    assert w32.dtype == np.float32
    assert np.allclose(w32, w64, rtol=1e-6)
    # sum-to-zero fix (fornberg.jl:57-61)
    assert abs(w64.sum()) <= 1e-15
    # (1,4): centre weight is the negated residual, not exactly 0
    w = O.calculate_weights(1, 0.0, np.arange(-2, 3))
    assert abs(w[2]) < 1e-15


def test_julia_cumsum_is_cumsum_to_rounding():
    # Base.cumsum is accumulate_pairwise!: c[i] = v1 + (v2 + ... + vi), pairwise above 128 elements;
    # equal to the sequential sum up to rounding only.
    for n in (3, 100, 1000):
        v = np.random.default_rng(0).random(n)
        assert np.allclose(O.julia_cumsum(v), np.cumsum(v), rtol=1e-13)
    assert np.array_equal(O.julia_cumsum(np.array([1.0, 2.0, 3.0])), [1.0, 3.0, 6.0])


def test_periodic_second_order_value_check():
    """test/DerivativeOperators/2nd_order_check.jl:3-38: L2*Q*u0 against the explicit periodic tridiagonal matrix, and
    L1*Q*(L1*Q*u0) against the explicit wide second-difference matrix (PeriodicBC, bc_operators.jl:17-19,:192)."""
    from oracle import oracle as O
    order, h = 2, 0.025 * np.pi
    N = int(2 * (np.pi / h))
    x = -np.pi + h * np.arange(N)
    L1 = O.CenteredDifference(1, order, h, N)
    L2 = O.CenteredDifference(2, order, h, N)
    Q = O.PeriodicBC(np.float64)
    u0 = np.cos(x)
    M = (np.diag(-2.0 * np.ones(N)) + np.diag(np.ones(N - 1), 1) + np.diag(np.ones(N - 1), -1))
    M[-1, 0] = 1.0
    M[0, -1] = 1.0
    M /= h ** 2
    np.testing.assert_allclose(O.apply_axis(L2, u0, Q), M @ u0, rtol=1e-9, atol=1e-9)
    A = np.zeros((N, N))
    for i in range(N):
        A[i, i] = -0.5
        for j in (i - 2, i + 2):
            A[i, j % N] = 0.25
    A /= h ** 2
    np.testing.assert_allclose(O.apply_axis(L1, O.apply_axis(L1, u0, Q), Q), A @ u0, rtol=1e-9, atol=1e-9)


def test_composite_equals_sum_of_axis_applications():
    """test/DerivativeOperators/2D_3D_fast_multiplication.jl (e.g. :24, :73, :159-161, :667-669): a summed composite
    applied to an array equals the sum of the per-axis applications; `overwrite=false` accumulates (convolutions.jl:49)."""
    from oracle import oracle as O
    rng = np.random.default_rng(3)
    shape = (13, 11, 12)
    u = np.asfortranarray(rng.uniform(-1, 1, shape))
    ops = [O.CenteredDifference(2, 4, 0.1 * (a + 1), shape[a], axis=a + 1) for a in range(3)] + \
          [O.UpwindDifference(1, 2, 0.05, shape[1], np.sin(np.arange(shape[1])), axis=2)]
    bcs = {a + 1: O.RobinBC((1.0, 2.0, 3.0), (0.0, -1.0, 2.0), 0.1 * (a + 1), 2) for a in range(3)}
    total = O.apply_sum(ops, u, bcs)
    parts = [O.apply_axis(op, u, bcs[op.axis]) for op in ops]
    acc = parts[0]
    for p in parts[1:]:
        acc = acc + p
    np.testing.assert_array_equal(total, acc)
    out = parts[0].copy(order="F")
    O.apply_axis(ops[1], u, bcs[2], out=out, overwrite=False)
    np.testing.assert_array_equal(out, parts[0] + parts[1])


def _matrix_by_multiplication(op, n):
    """Array(L) by applying the 1-D operator to unit vectors of the padded space (concretization.jl:8-48 equivalent)."""
    from oracle import oracle as O
    M = np.zeros((n, n + 2))
    for k in range(n + 2):
        e = np.zeros(n + 2)
        e[k] = 1.0
        M[:, k] = O.apply_axis(op, e)
    return M


def test_nd_nonsymmetric_stencils_and_coefficients_equal_matrix_times_pencils():
    """test/DerivativeOperators/differentiation_dimension.jl:209-302 (non-symmetric interior stencil (3,4)) and :304-370
    (operators with coefficients): mul! along axis N of a 2-D / 3-D array equals Array(L) applied to every pencil."""
    from oracle import oracle as O
    n = 30
    for coeff in (1, 2.5, np.linspace(0.5, 1.5, n)):
        for shape, axis in [((32, 32), 1), ((32, 32), 2), ((32, 32, 32), 1), ((32, 32, 32), 2), ((32, 32, 32), 3)]:
            L = O.CenteredDifference(3, 4, 0.1, n, coeff, axis=axis)
            A = _matrix_by_multiplication(O.CenteredDifference(3, 4, 0.1, n, coeff, axis=1), n)
            idx = np.indices(shape)[axis - 1] + 1
            M = np.asfortranarray(np.cos(0.1 * idx))
            got = O.apply_axis(L, M)
            want = np.moveaxis(np.tensordot(A, np.moveaxis(M, axis - 1, 0), axes=(1, 0)), 0, axis - 1)
            assert got.shape == tuple(s - 2 if d == axis - 1 else s for d, s in enumerate(shape))
            np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-9)


def test_regular_operator_validation():
    """test/DerivativeOperators/regular_operator_validation.jl:3-17: derivative orders 1..6 and approximation orders 2..10 of
    sin(x) on 200 points; the reference's own assertion is isapprox(...; atol = 1.0^(1 - aor)) = a 2-norm of the error of at
    most 1 (kept literally), and the interior rows have to be accurate to the truncation order as well."""
    from oracle import oracle as O
    nx = 200
    x = np.linspace(0, 2 * np.pi, nx)
    dx = x[1] - x[0]
    y = np.sin(x)
    dy = [np.cos(x), -np.sin(x), -np.cos(x), y, np.cos(x), -np.sin(x)]
    for dor in range(1, 7):
        for aor in range(2, 11, 2):
            D1 = O.CenteredDifference(dor, aor, dx, nx - 2)
            dyt = O.apply_axis(D1, y)
            err = dyt - dy[dor - 1][1:-1]
            assert np.linalg.norm(err) <= 1.0, (dor, aor, np.linalg.norm(err))
            r = D1.stencil_length // 2 + 1
            assert np.abs(err[r:-r]).max() <= 20 * dx ** aor + 1e-15 / dx ** dor * 2e3, (dor, aor, np.abs(err[r:-r]).max())


def test_generic_operator_validation():
    """test/DerivativeOperators/generic_operator_validation.jl:3-60: the irregular-grid constructor fed a regular grid
    differentiates sin(2x) to atol 10^(1-aor) (2-norm), and on a genuinely irregular grid to the reference's tolerance
    2 * 10^(2-aor) * max(dx)^(2-dor)."""
    from oracle import oracle as O
    x = np.arange(0.0, np.pi, 0.01)
    dxv = np.diff(x)
    y = np.sin(2 * x)
    x_, y_ = x[1:-1], y[1:-1]
    dy = [2 * np.cos(2 * x_), -4 * np.sin(2 * x_), -8 * np.cos(2 * x_), 16 * y_, 32 * np.cos(2 * x_)]
    for dor in range(1, 5):
        for aor in range(2, 7):
            D1 = O.CenteredDifference(dor, aor, dxv, len(x) - 2)
            assert np.linalg.norm(O.apply_axis(D1, y) - dy[dor - 1]) <= 10.0 ** (1 - aor), (dor, aor)
            # ... and its rows equal the regular-grid operator's (sparse(Dr) ~ sparse(Dir), :11-19)
            if aor % 2 == 0:
                Dr = O.CenteredDifference(dor, aor, dxv[0], len(x) - 2)
                np.testing.assert_allclose(O.apply_axis(D1, y), O.apply_axis(Dr, y), rtol=0, atol=1e-6 * np.abs(dy[dor - 1]).max() / dxv[0] ** 0)
    x = np.cumsum(np.sin(np.arange(0.0, np.pi, 0.05)))
    x = x / x[-1] * np.pi
    dxv = np.diff(x)
    y = np.sin(2 * x)
    x_, y_ = x[1:-1], y[1:-1]
    dy = [2 * np.cos(2 * x_), -4 * np.sin(2 * x_), -8 * np.cos(2 * x_), 16 * y_, 32 * np.cos(2 * x_)]
    for dor in range(1, 5):
        for aor in range(4, 11):
            D1 = O.CenteredDifference(dor, aor, dxv, len(x) - 2)
            tol = 2 * 10.0 ** (2 - aor) * dxv.max() ** (2 - dor)
            assert np.linalg.norm(O.apply_axis(D1, y) - dy[dor - 1]) <= tol, (dor, aor, tol)


def test_heat_equation_with_dirichlet_and_neumann_bcs():
    """test/DerivativeOperators/heat_eqn.jl:4-30, :32-66: du/dt = A*bc*u integrated over [0, 1] with an explicit 5th-order
    Runge-Kutta pair (the reference uses Tsit5; scipy's RK45 here) keeps the end values (Dirichlet, rtol 0.05) and the end
    slopes (Neumann, atol 0.1); the (2,1) upwind operator with one offside point behaves like the centered one."""
    from scipy.integrate import solve_ivp
    from oracle import oracle as O
    times = np.arange(0.0, 1.01, 0.1)
    # Dirichlet
    h = 2 * np.pi / 511
    x = -np.pi + h * np.arange(512)
    ua = lambda z: -(z - 0.5) ** 2 + 1 / 12
    u0 = ua(x)
    bc = O.DirichletBC(ua(-np.pi - h), ua(np.pi + h))
    for A in (O.CenteredDifference(2, 2, h, 512), O.UpwindDifference(2, 1, h, 512, 1, offside=1)):
        sol = solve_ivp(lambda t, u: O.apply_axis(A, u, bc), (0.0, 1.0), u0, method="RK45", rtol=1e-3, atol=1e-6, dense_output=True)
        assert sol.status == 0
        for t in times:
            s = sol.sol(t)
            assert abs(s[0] - u0[0]) <= 0.05 * abs(u0[0]) and abs(s[-1] - u0[-1]) <= 0.05 * abs(u0[-1]), t
    # Neumann
    N = 128
    dx = 2 * np.pi / (N - 1)
    x = -np.pi + dx * np.arange(N)
    u0 = ua(x)
    c0 = np.array([-11 / 6, 3.0, -3 / 2, 1 / 3]) / dx
    c1 = -c0[::-1]
    for B, A in ((O.CenteredDifference(1, 2, dx, N - 2), O.CenteredDifference(2, 2, dx, N)),
                 (O.UpwindDifference(1, 2, dx, N - 2, 1, offside=1), O.UpwindDifference(2, 1, dx, N, 1, offside=1))):
        d = O.apply_axis(B, u0)
        bcn = O.NeumannBC((d[0], d[-1]), dx, 1)
        sol = solve_ivp(lambda t, u: O.apply_axis(A, u, bcn), (0.0, 1.0), u0, method="RK45", rtol=1e-3, atol=1e-6, dense_output=True)
        assert sol.status == 0
        for t in times:
            s = sol.sol(t)
            assert abs(c0 @ s[:4] - d[0]) <= 1e-1 and abs(c1 @ s[-4:] - d[-1]) <= 1e-1, t


def test_fornberg_weights_against_the_reference_stencil_literals():
    """Literal interior stencils held by the reference's tests (derivative_operators_interface.jl:146-227: "Correctness of
    Uniform Stencils, Complete", "... Complete Half", "... Uniform Upwind Stencils"), each the Fornberg weights of the point set
    its constructor passes to calculate_weights (derivative_operator.jl:337-339 half-offset x0 = 0.5, :655-658 upwind):
    pins fornberg.jl:7-62 in the oracle and in the host mirror beyond the operator matrices."""
    import deo_b200 as D
    from oracle import oracle as O
    cases = []
    centered = {2: ([-0.5, 0, 0.5], [1.0, -2.0, 1.0], [-1 / 2, 1.0, 0.0, -1.0, 1 / 2]),
                4: ([1 / 12, -2 / 3, 0, 2 / 3, -1 / 12], [-1 / 12, 4 / 3, -5 / 2, 4 / 3, -1 / 12], [1 / 8, -1.0, 13 / 8, 0.0, -13 / 8, 1.0, -1 / 8])}
    for a, ws in centered.items():
        for d, w in enumerate(ws, start=1):
            sl = d + a - 1 + (d + a) % 2
            cases.append((d, 0.0, np.arange(-(sl // 2), sl // 2 + 1, dtype=float), w))
    half = {2: ([0.5, 0.5], [-1.0, 1.0]), 4: ([-1 / 16, 9 / 16, 9 / 16, -1 / 16], [1 / 24, -9 / 8, 9 / 8, -1 / 24])}
    for a, ws in half.items():
        for d, w in enumerate(ws):
            sl = a + 2 * (d // 2) + a % 2
            end = sl // 2
            cases.append((d, 0.5, np.arange(1 - end, end + 1, dtype=float), w))
    upwind = {1: ([-1.0, 1.0], [-1.0, 3.0, -3.0, 1.0]), 2: ([-3 / 2, 2.0, -1 / 2], [-5 / 2, 9.0, -12.0, 7.0, -3 / 2])}
    for a, ws in upwind.items():
        for d, w in zip((1, 3), ws):
            cases.append((d, 0.0, np.arange(0.0, d + a), w))
    assert len(cases) == 14
    for d, x0, x, want in cases:
        for got in (O.calculate_weights(d, x0, x), D.calculate_weights(d, x0, x)):
            np.testing.assert_allclose(np.asarray(got, dtype=float), want, rtol=0, atol=1e-10, err_msg=f"d={d} x0={x0} x={x}")


def test_operations_on_matrices():
    """test/DerivativeOperators/derivative_operators_interface.jl:361-405: a (2,2) centered operator applied to matrices column
    by column, uniform and non-uniform grids, against the analytic second derivatives of x^2 + y and x^2 + y^2 with the
    reference's tolerances."""
    from oracle import oracle as O
    N, M = 51, 101
    x1, y1 = np.linspace(0, 1, N), np.linspace(0, 1, M)
    x2 = np.concatenate([x1[: N // 2] ** 0.2020, x1[(N + 1) // 2 - 1:] ** 2.015])
    y2 = np.concatenate([y1[: M // 2] ** 1.793, y1[(M + 1) // 2 - 1:] ** 2.019])
    for i, (xa, ya) in enumerate(((x1, y1), (x2, y2)), start=1):
        dx, dy = (xa[1] - xa[0], ya[1] - ya[0]) if i == 1 else (np.diff(xa), np.diff(ya))
        F = np.asfortranarray(xa[:, None] ** 2 + ya[None, :])
        A = O.CenteredDifference(2, 2, dx, len(xa) - 2)
        AF = O.apply_axis(A, F)
        assert AF.shape == (len(xa) - 2, len(ya))
        assert np.abs(AF - 2.0).max() <= 1e-9
        if i == 2:      # (for i == 1 the reference builds B with len(yarr) rows for an input of len(yarr) rows -- a length mismatch, not reproduced)
            B = O.CenteredDifference(2, 2, dy, len(ya) - 2)
            BFt = O.apply_axis(B, np.asfortranarray(F.T))
            assert np.abs(BFt).max() <= 1e-8
            assert np.abs(O.apply_axis(A, np.asfortranarray(BFt.T))).max() <= 1e-4
            G = np.asfortranarray(xa[:, None] ** 2 + ya[None, :] ** 2)
            assert np.abs(O.apply_axis(A, G) - 2.0).max() <= 1e-9
            BGt = O.apply_axis(B, np.asfortranarray(G.T))
            assert BGt.shape == (len(ya) - 2, len(xa)) and np.abs(BGt - 2.0).max() <= 1e-8
            assert np.abs(O.apply_axis(A, np.asfortranarray(BGt.T))).max() <= 1e-4


def test_heat_equation_example_against_the_analytic_solution():
    """test/DerivativeOperators/heat_equation.jl:11-36: du/dt = Delta*bc*u, Dirichlet0, u0 = sin(2 pi x) on 100 knots; at
    t = 0.03 the solution equals sin(2 pi x) exp(-t (2 pi)^2) to rtol 1e-3 (the reference integrates with KenCarp4)."""
    from scipy.integrate import solve_ivp
    from oracle import oracle as O
    nknots = 100
    h = 1.0 / (nknots + 1)
    knots = h * np.arange(1, nknots + 1)
    ua = lambda x, t: np.sin(2 * np.pi * x) * np.exp(-t * (2 * np.pi) ** 2)
    L = O.CenteredDifference(2, 2, h, nknots)
    bc = O.Dirichlet0BC(np.float64)
    sol = solve_ivp(lambda t, u: O.apply_axis(L, u, bc), (0.0, 0.03), ua(knots, 0.0), method="RK45", rtol=1e-6, atol=1e-9)
    assert sol.status == 0
    want = ua(knots, 0.03)
    assert np.linalg.norm(sol.y[:, -1] - want) <= 1e-3 * max(np.linalg.norm(want), np.linalg.norm(sol.y[:, -1]))


def test_kdv_single_soliton_advection_with_negative_upwind_coefficient():
    """test/DerivativeOperators/KdV.jl:4-73: du = A*bc*u with A = UpwindDifference(1, 3, dx, n, -1) -- the c < 0 branch of the
    upwind convolutions (convolutions.jl:134-139, :181-210) -- a time-dependent five-coefficient GeneralBC, plain, with one
    offside point, and through the non-uniform constructor on a uniform grid; the travelling wave sech^2((x - t)/2)/2 is
    reproduced to a 2-norm of 0.01 at t = 0, 0.5, ..., 5 (Tsit5 with tolerances 1e-6 in the reference; RK45 here)."""
    from scipy.integrate import solve_ivp
    from oracle import oracle as O
    dx = 1.0 / 20
    x = -10 + dx * np.arange(401)
    phi = lambda z, t: 0.5 / np.cosh((z - t) / 2) ** 2
    n = len(x)
    ops = [O.UpwindDifference(1, 3, dx, n, -1.0), O.UpwindDifference(1, 3, dx, n, -1.0, offside=1),
           O.UpwindDifference(1, 3, dx * np.ones(n + 1), n, -1.0, offside=1)]
    for A in ops:
        def rhs(t, u, A=A):
            bc = O.GeneralBC([0, 1, -6 * phi(-10, t), 0, -1], [0, 1, -6 * phi(10, t), 0, -1], dx, 3)
            return O.apply_axis(A, u, bc)
        sol = solve_ivp(rhs, (0.0, 5.0), phi(x, 0.0), method="RK45", rtol=1e-6, atol=1e-6, dense_output=True)
        assert sol.status == 0
        for t in np.arange(0.0, 5.01, 0.5):
            assert np.linalg.norm(sol.sol(t) - phi(x, t)) <= 0.01, (t, np.linalg.norm(sol.sol(t) - phi(x, t)))
