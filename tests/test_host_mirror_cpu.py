"""Host-side mirror of the reference's operator / BC interface (diffeqoperators.jl_b200/operators.py, bc.py): the parts
that need no device -- constructor rules, composition, coefficient scaling, error behaviour -- against the reference's
documented behaviour (file:line cited per check) and against the oracle's restatement."""
import numpy as np
import pytest


@pytest.fixture(scope="module")
def D():
    import deo_b200
    return deo_b200


def test_stencil_geometry_rules(D):
    # derivative_operator.jl:85-92 (centered), :461-463 (upwind)
    for d, a in [(1, 2), (2, 2), (2, 4), (2, 6), (3, 2), (4, 4)]:
        A = D.CenteredDifference(d, a, 0.1, 40)
        sl = d + a - 1 + (d + a) % 2
        assert (A.stencil_length, A.boundary_stencil_length, A.boundary_point_count) == (sl, d + a, sl // 2 - 1)
        assert A.size == (40, 42)
    for d, a, off in [(1, 1, 0), (1, 2, 0), (2, 3, 1), (1, 3, 1)]:
        A = D.UpwindDifference(d, a, 0.1, 40, 1.0, offside=off)
        assert (A.stencil_length, A.boundary_stencil_length, A.boundary_point_count, A.offside) == (d + a, d + a, d + a - 2 - off, off)
    with pytest.raises(AssertionError):
        D.UpwindDifference(1, 2, 0.1, 40, 1.0, offside=-1)            # :458
    with pytest.raises(AssertionError):
        D.UpwindDifference(1, 2, 0.1, 40, 1.0, offside=2)             # :459


def test_scalar_and_vector_scaling_follow_the_reference(D):
    # derivative_operator_functions.jl:165-197: c*A scales `coefficients`, the stencils are untouched
    A = D.CenteredDifference(2, 2, 0.5, 8)
    B = 3.3 * A
    assert np.array_equal(B.stencil_coefs, A.stencil_coefs) and np.allclose(B.coefficients, 3.3)
    c = np.linspace(1, 2, 8)
    C = c * A
    assert np.array_equal(C.coefficients, c * A.coefficients)
    with pytest.raises(ValueError, match="DimensionMismatch"):          # :182-184
        np.ones(7) * A
    assert np.allclose((-A).coefficients, -1.0)


def test_coefficient_functions(D):
    # coefficient_functions.jl:7-26: number, vector, or function of the (zero-initialised) coefficient vector
    assert np.allclose(D.CenteredDifference(2, 2, 0.1, 6, 2.5).coefficients, 2.5)
    v = np.arange(6.0)
    assert np.array_equal(D.CenteredDifference(2, 2, 0.1, 6, v).coefficients, v)
    A = D.CenteredDifference(2, 2, 0.1, 6, lambda x: x + 1.5)
    assert np.allclose(A.coefficients, 1.5)
    with pytest.raises(ValueError):
        D.CenteredDifference(2, 2, 0.1, 6, np.ones(5))
    # update_coefficients!(A,u,p,t) mutates in place and invalidates cached plans (abstract_operator_functions.jl:190-194)
    def cf(coeffs, u, p, t):
        coeffs[:] = p * t
    U = D.CenteredDifference(2, 2, 0.1, 6, 1.0)
    U.coeff_func = cf
    v0 = U._version
    U.update_coefficients_(None, 2.0, 3.0)
    assert np.allclose(U.coefficients, 6.0) and U._version == v0 + 1


def test_composition_structure(D):
    # composite_operators.jl:15-46 (sum keeps its operands), ghost_derivative_operator.jl:7-13 (L*Q, distribution over sums)
    shape = (10, 12, 14)
    ops = [D.CenteredDifference[ax](2, 2, 0.1, shape[ax - 1]) for ax in (1, 2, 3)]
    S = ops[0] + ops[1] + ops[2]
    assert isinstance(S, D.DiffEqOperatorCombination) and len(S.ops) == 3
    Q = D.compose(*D.Dirichlet0BC(np.float64, shape))
    G = S * Q
    from deo_b200.apply import _terms
    assert [(L.axis, Qk is Q) for L, Qk in _terms(G)] == [(1, True), (2, True), (3, True)]
    G1 = ops[1] * Q
    assert isinstance(G1, D.GhostDerivativeOperator) and G1.L is ops[1] and G1.Q is Q
    with pytest.raises(TypeError):
        D.CenteredDifference(2, 2, 0.1, 10, dtype=np.float32) * D.Dirichlet0BC(np.float64)   # eltype mismatch


def test_compose_errors(D):
    # multi_dim_bc_operators.jl:178-192
    shape = (6, 7, 8)
    Qx, Qy, Qz = D.RobinBC((1.0, 0.5, 0.25), (1.0, -0.5, 0.75), (0.1, 0.1, 0.1), 1, shape)
    assert isinstance(D.compose(Qx, Qy, Qz), D.ComposedMultiDimBC)
    with pytest.raises(ValueError, match="enough BCs"):
        D.compose(Qx, Qy)
    with pytest.raises(ValueError, match="same dimension"):
        D.compose(Qx, Qx, Qz)


def test_bc_stencils_match_the_oracle(D):
    # bc_operators.jl:26-39 (Robin), :85-170 (General), :173-183 (Neumann / Dirichlet special cases)
    from oracle import oracle as O
    for T in (np.float64, np.float32):
        for order in (1, 2, 3, 4):
            p, o = D.RobinBC((0.7, 1.3, -0.4), (1.1, -0.6, 2.0), 0.05, order, dtype=T), O.RobinBC((0.7, 1.3, -0.4), (1.1, -0.6, 2.0), 0.05, order, T)
            assert np.array_equal(p.a_l, o.a_l) and np.array_equal(p.a_r, o.a_r) and p.b_l == o.b_l and p.b_r == o.b_r
            p, o = D.NeumannBC((0.3, -0.2), 0.05, order, dtype=T), O.NeumannBC((0.3, -0.2), 0.05, order, T)
            assert np.array_equal(p.a_l, o.a_l) and np.array_equal(p.a_r, o.a_r) and p.b_l == o.b_l and p.b_r == o.b_r
        al, ar = [1.0, 2.0, -0.5, 0.25], [0.5, -1.0, 0.75]
        p, o = D.GeneralBC(al, ar, 0.1, 3, dtype=T), O.GeneralBC(al, ar, 0.1, 3, T)
        assert np.allclose(p.a_l, o.a_l, rtol=0, atol=0) and np.allclose(p.a_r, o.a_r, rtol=0, atol=0) and p.b_l == o.b_l and p.b_r == o.b_r
    d = D.DirichletBC(2.0, 3.0)
    assert d.b_l == 2.0 and d.b_r == 3.0 and np.all(d.a_l == 0) and np.all(d.a_r == 0)       # 1/0 -> Inf -> a = [-0.0, 0.0]


def test_host_side_padded_array_matches_the_reference_rules(D):
    # bc_operators.jl:188-192, multi_dim_bc_operators.jl:212-252, boundary_padded_arrays.jl:158-180 (corners read as 0)
    from deo_b200.apply import materialize_padded
    rng = np.random.default_rng(0)
    u = rng.uniform(-1, 1, 9)
    Q = D.RobinBC((1.0, 6.0, 10.0), (1.0, 6.0, 10.0), 1.0, 3)
    P = materialize_padded(Q * u)
    assert P.shape == (11,) and np.array_equal(P[1:-1], u)
    assert np.isclose(P[0], Q.a_l @ u[:3] + Q.b_l) and np.isclose(P[-1], Q.a_r @ u[-3:] + Q.b_r)
    # robin.jl:92-100: u = 1:10, (1, 6, 10), dx = 1, order 3 -> ghosts -4/10 and 125/12
    P = materialize_padded(Q * np.arange(1.0, 11.0))
    assert np.isclose(P[0], -4 / 10) and np.isclose(P[-1], 125 / 12)
    A = rng.uniform(-1, 1, (5, 6))
    Qc = D.compose(*D.Dirichlet0BC(np.float64, A.shape))
    PA = materialize_padded(Qc * A)
    assert PA.shape == (7, 8) and PA[0, 0] == 0 and PA[-1, -1] == 0 and np.array_equal(PA[1:-1, 1:-1], A)


def test_multi_dim_bc_extension_follows_the_reference_test(D):
    """test/DerivativeOperators/multi_dim_bc_test.jl:9-35, :41-88, :91-108 (commented out in the reference's runtests.jl; the rules are those of
    multi_dim_bc_operators.jl:27-52, :212-252): MultiDimBC arrays that mix RobinBC and PeriodicBC
    pencils extend every pencil exactly as the atomic BC extends the vector (generic slice_rmul: the 1-D periodic rule),
    Dirichlet0 along the third axis, and compose(Q1..QN) * A equals the per-axis extensions for 2..6 dimensions."""
    from deo_b200.apply import materialize_padded
    rng = np.random.default_rng(7373)
    q1 = D.RobinBC((1.0, 2.0, 3.0), (0.0, -1.0, 2.0), 0.1, 4)
    q2 = D.PeriodicBC(np.float64)

    def mixed(face):
        arr = np.empty(face, dtype=object)
        first = face[0] // 2
        for idx in np.ndindex(*face):
            arr[idx] = q1 if idx[0] < first else q2
        return arr

    # 2-D
    n, m = 8, 15
    A = rng.random((n, m))
    BCx, BCy = mixed((m,)), mixed((n,))
    Ax = materialize_padded(D.MultiDimBC[1](BCx) * A)
    Ay = materialize_padded(D.MultiDimBC[2](BCy) * A)
    assert Ax.shape == (n + 2, m) and Ay.shape == (n, m + 2)
    for j in range(m):
        assert np.array_equal(Ax[:, j], materialize_padded(BCx[j] * A[:, j]))
    for i in range(n):
        assert np.array_equal(Ay[i, :], materialize_padded(BCy[i] * A[i, :]))
    # 3-D
    n, m, o = 8, 11, 12
    A = rng.random((n, m, o))
    BCx, BCy = mixed((m, o)), mixed((n, o))
    Qz = D.MultiDimBC[3](D.Dirichlet0BC(np.float64), A.shape)
    Ax = materialize_padded(D.MultiDimBC[1](BCx) * A)
    Ay = materialize_padded(D.MultiDimBC[2](BCy) * A)
    Az = materialize_padded(Qz * A)
    assert Ax.shape == (n + 2, m, o) and Ay.shape == (n, m + 2, o) and Az.shape == (n, m, o + 2)
    for j in range(m):
        for k in range(o):
            assert np.array_equal(Ax[:, j, k], materialize_padded(BCx[j, k] * A[:, j, k]))
    for i in range(n):
        for k in range(o):
            assert np.array_equal(Ay[i, :, k], materialize_padded(BCy[i, k] * A[i, :, k]))
    z0 = D.Dirichlet0BC(np.float64)
    for i in range(n):
        for j in range(m):
            assert np.array_equal(Az[i, j, :], materialize_padded(z0 * A[i, j, :]))
    # an array whose every pencil is periodic is the reference's specialised method: ghosts swapped (:221-228)
    allp = np.empty((m, o), dtype=object)
    for idx in np.ndindex(m, o):
        allp[idx] = q2
    Ap = materialize_padded(D.MultiDimBC[1](allp) * A)
    assert np.array_equal(Ap[0], A[0]) and np.array_equal(Ap[-1], A[-1])
    # compositions to higher dimension
    for N in range(2, 7):
        sizes = tuple(int(v) for v in rng.integers(4, 8, N))
        A = rng.random(sizes)
        Q1_N = D.RobinBC(tuple(rng.random(3)), tuple(rng.random(3)), [0.1] * N, 4, sizes)
        full = materialize_padded(D.compose(*Q1_N) * A)
        assert full.shape == tuple(s + 2 for s in sizes)
        for d, Qd in enumerate(Q1_N):
            one = materialize_padded(Qd * A)                    # padded along axis d only
            sel = tuple(slice(None) if e == d else slice(1, -1) for e in range(N))
            assert np.array_equal(full[sel], one)
