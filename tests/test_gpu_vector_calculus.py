"""SURVEY 8(f)-3 / 8(f)-4 on the device: Gradient / Divergence / Curl (vector_calculus_operators.jl), nonlinear_diffusion!
(derivative_operator.jl:31-77) and DiffEqOperatorComposition (composite_operators.jl:86-212), every derivative one
libdeo_b200 plan.  Checked against the oracle's 1-D pieces on random (non-separable) fields, and against the reference's own
analytic tests (test/DerivativeOperators/gradient_operator.jl, divergence_operator.jl, curl_operator.jl)."""
from math import comb

import numpy as np
import pytest

from tests.helpers import TOL, assert_close, make_pair, uniform_field

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def D():
    import deo_b200
    deo_b200.load_library()
    return deo_b200


@pytest.fixture(scope="module")
def O():
    from oracle import oracle
    return oracle


@pytest.fixture(scope="module")
def golden():
    import json, os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_kats.json")) as f:
        return json.load(f)


def _d1(O, a, dx, n, axis, dtype, scale=1):
    op = O.CenteredDifference(1, a, dx, n, axis=axis, dtype=dtype)
    return op.scale(scale) if scale != 1 else op


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_gradient_divergence_curl_match_the_oracle_on_random_fields(D, O, dtype):
    lens, a = (14, 14, 14), 4
    dx = (0.1, 0.2, 0.3)
    pshape = tuple(l + 2 for l in lens)
    M = uniform_field(pshape, dtype, seed=1)
    V = uniform_field(pshape + (3,), dtype, seed=2)
    zeros = lambda: np.zeros(lens, dtype=dtype, order="F")
    # Gradient: out[.., n] = d/dx_n M, perpendicular padding stripped
    g = D.Gradient(a, dx, lens, dtype=dtype) * M
    for n in range(3):
        assert_close(g[..., n], O.apply_axis(_d1(O, a, dx[n], lens[n], n + 1, dtype), M, None, out=zeros()), dtype, f"gradient component {n + 1}")
    # Divergence: sum_n d/dx_n V[.., n]
    want = sum(O.apply_axis(_d1(O, a, dx[n], lens[n], n + 1, dtype), np.asfortranarray(V[..., n]), None, out=zeros()) for n in range(3))
    assert_close(D.Divergence(a, dx, lens, dtype=dtype) * V, want, dtype, "divergence")
    # Curl
    c = D.Curl(a, dx, lens, dtype=dtype) * V
    d = lambda n, comp: O.apply_axis(_d1(O, a, dx[n], lens[n], n + 1, dtype), np.asfortranarray(V[..., comp]), None, out=zeros())
    ref = [d(1, 2) - d(2, 1), d(2, 0) - d(0, 2), d(0, 1) - d(1, 0)]
    scale = max(np.abs(r).max() for r in ref)
    for n in range(3):
        assert np.abs(c[..., n].astype(np.float64) - ref[n]).max() <= TOL[np.dtype(dtype)] * scale, f"curl component {n + 1}"
    # c * A (vector_calculus_operators.jl:75-88)
    assert_close((2 * D.Curl(a, dx, lens, dtype=dtype)) * V, 2 * c, dtype, "2 * Curl")


def test_reference_analytic_cases(D):
    """gradient_operator.jl:5-24, divergence_operator.jl, curl_operator.jl:3-32: quadratic fields, uniform and vector dx."""
    x = np.arange(-5, 5.0001, 1.25)
    n = x.size
    X, Y, Z = np.meshgrid(x, x, x, indexing="ij")
    for dx in (1.25, 1.25 * np.ones(n - 1)):
        dxs = (dx, dx, dx)
        lens = (n - 2,) * 3
        grad = D.Gradient(4, dxs, lens) * np.asfortranarray(X ** 2 + Y ** 2 + Z ** 2)
        for c, W in enumerate((X, Y, Z)):
            assert np.abs(grad[..., c] - 2 * W[1:-1, 1:-1, 1:-1]).max() < 1e-3
        u0 = np.zeros((n, n, n, 3), order="F")
        u0[..., 0], u0[..., 1], u0[..., 2] = Y ** 2 + Z ** 2, X ** 2 + Z ** 2, X ** 2 + Y ** 2
        curl = D.Curl(4, dxs, lens) * u0
        inner = (slice(1, -1),) * 3
        for c, W in enumerate((2 * (Y - Z), 2 * (Z - X), 2 * (X - Y))):
            assert np.abs(curl[..., c] - W[inner]).max() < 1e-3
        v0 = np.zeros((n, n, n, 3), order="F")
        v0[..., 0], v0[..., 1], v0[..., 2] = X ** 2, Y ** 2, Z ** 2
        div = D.Divergence(4, dxs, lens) * v0
        assert np.abs(div - 2 * (X + Y + Z)[inner]).max() < 1e-3


@pytest.mark.parametrize("s,f", [(1, 0), (1, 1), (2, 0), (2, 1), (3, 1)])
def test_nonlinear_diffusion(D, O, s, f):
    """du = sum of products of derivatives of the padded vectors q (= bc*u) and p (derivative_operator.jl:31-66)."""
    n, a, dx = 120, 4, 0.05
    q = uniform_field(n + 2, np.float64, seed=3)
    p = uniform_field(n + 2, np.float64, seed=4)
    Dq = lambda k: O.apply_axis(O.CenteredDifference(k, a, dx, n), q)
    Dp = lambda k: O.apply_axis(O.CenteredDifference(k, a, dx, n), p)
    want = (Dq(f) if f > 0 else q[1:-1]) * Dp(s)
    for l in range(1, s):
        want = want + comb(s, l) * Dq(l + f) * Dp(s - l)
    want = want + Dq(f + s) * p[1:-1]
    got = D.nonlinear_diffusion(s, f, a, p, q, dx, n)
    assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max()
    du = np.zeros(n)
    D.nonlinear_diffusion_(du, s, f, a, p, q, dx, n)
    assert np.array_equal(du, got)


def test_operator_composition(D, O):
    """(L2*Q2) ∘ (L1*Q1): apply L1*Q1, then L2*Q2 (composite_operators.jl:113-115, :150-152, :186-192)."""
    shape = (40, 36, 32)
    h = (0.1, 0.1, 0.1)
    u = uniform_field(shape, np.float64, seed=5)
    robin = ((1.0, 0.5, 0.25), (1.0, -0.5, 0.75))
    Q = D.compose(*D.RobinBC(*robin, h, 1, shape))
    bcs = {ax + 1: O.RobinBC(*robin, h[ax], 1) for ax in range(3)}
    pairs = [make_pair("centered", 2, 4, h[ax], shape[ax], axis=ax + 1) for ax in range(3)]
    lap = (pairs[0][0] + pairs[1][0] + pairs[2][0]) * Q
    gx = D.CenteredDifference[1](1, 4, h[0], shape[0]) * Q
    comp = D.compose_operators(gx, lap)                       # d/dx after the Laplacian
    step1 = O.apply_sum([pr[1] for pr in pairs], u, bcs)
    want = O.apply_axis(O.CenteredDifference(1, 4, h[0], shape[0], axis=1), step1, bcs[1])
    ud = D.DeviceArray.from_host(u)
    assert_close((comp * ud).to_host(), want, np.float64, "composition *")
    y = D.DeviceArray(shape, np.float64)
    comp.mul_(y, ud)
    comp.mul_(y, ud)                                          # second call runs through the cached intermediate
    assert_close(y.to_host(), want, np.float64, "composition mul!")
    both = gx * lap                                           # `*` of two operators builds the same composition (:107-112)
    assert isinstance(both, D.DiffEqOperatorComposition) and both.ops == comp.ops


def test_concretization_and_ldiv(D, golden):
    """poisson.jl:6-35: (Δ*bc) \\ fill(f, n) reproduces the quadratic analytic solution; the concretized L*Q equals the
    reference's golden 3x3 matrices (BasicSDOExamples.jl:49-59 via tests/golden) and a 2-D solve satisfies A x = b."""
    f, a, b, n = 1.0, -1.0, 2.0, 10
    h = 1.0 / (n + 1)
    G = D.CenteredDifference(2, 2, h, n) * D.DirichletBC(a, b)
    u = D.ldiv(G, np.full(n, f))
    x = h * np.arange(1, n + 1)
    assert np.allclose(u, f / 2 * x ** 2 + (b - a - f / 2) * x + a, rtol=1e-12, atol=1e-12)
    g = golden["ghost_operator_matrices"]
    for case in g["cases"]:
        if len(case["terms"]) != 1 or "scale" in case["terms"][0]:
            continue
        t = case["terms"][0]
        Q = D.Neumann0BC(g["dx"], 1) if case["bc"]["type"] == "neumann0" else D.RobinBC(case["bc"]["l"], case["bc"]["r"], g["dx"], case["bc"]["order"])
        mk = D.CenteredDifference if t["kind"] == "centered" else D.UpwindDifference
        A_l, _ = D.concretize(mk(t["d"], t["a"], g["dx"], g["M"], t["coeff"]) * Q, g["M"])
        want = np.asarray(case["matrix"]) / case.get("scale", 1.0)
        assert np.abs(A_l - want).max() <= 1e-9 * np.abs(want).max() + 1e-12, case["name"]
    # 2-D: residual of the solve through the fused application itself
    shape = (12, 9)
    hh = (0.1, 0.15)
    A = D.CenteredDifference[1](2, 2, hh[0], shape[0]) + D.CenteredDifference[2](2, 2, hh[1], shape[1])
    G2 = A * D.compose(*D.RobinBC((1.0, 0.5, 0.25), (1.0, -0.5, 0.75), hh, 1, shape))
    rhs = uniform_field(shape, np.float64, seed=9)
    sol = D.ldiv(G2, rhs)
    assert np.abs(G2 * sol - rhs).max() <= 1e-9 * np.abs(rhs).max()
