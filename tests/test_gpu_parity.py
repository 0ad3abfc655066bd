"""GPU parity tests: every path goes product host mirror -> C ABI (libdeo_b200.so) -> CUDA kernel and is
compared with the CPU oracle on the same seeded inputs.  Tolerances are north_star's: <= 1e-13 relative
(Float64), <= 1e-5 (Float32), max over all points including boundaries, relative to max|oracle|."""
import itertools

import numpy as np
import pytest

from tests.helpers import TOL, assert_close, bc_pair, make_pair, nonuniform_dx, rel_err, uniform_field

pytestmark = pytest.mark.gpu

DTYPES = [np.float64, np.float32]
CENTERED = [(1, 2), (2, 2), (1, 4), (2, 4), (2, 6), (4, 4), (3, 2), (4, 10), (8, 8)]
UPWIND = [(1, 1, 0), (1, 2, 0), (2, 2, 0), (2, 3, 0), (1, 3, 1), (2, 3, 2), (3, 3, 1), (1, 4, 1), (1, 5, 2)]
BCS = [("dirichlet0",), ("dirichlet", 0.7, -1.3), ("neumann", (0.3, -0.2), 1), ("robin", (1.0, 0.5, 0.25), (1.0, -0.5, 0.75), 1),
       ("robin", (1.0, 6.0, 10.0), (1.0, 6.0, 10.0), 3), ("general", (-10.0, 1.0, 6.0), (-2.0, 1.5, 3.0), 3),
       ("general", (0.5, 1.0, 6.0, 0.2), (-2.0, 1.5, 3.0, 0.1), 2), ("periodic",)]


@pytest.fixture(scope="module")
def D():
    import deo_b200
    deo_b200.load_library()
    return deo_b200


@pytest.fixture(scope="module")
def O():
    from oracle import oracle
    return oracle


def _flags(D, generic):
    return D._lib.DEO_FLAG_FORCE_GENERIC if generic else 0


# ------------------------------------------------------------------------------------------------------
# 1-D, plain padded vector:  mul!(y, A, x)   (convolutions.jl:17-22, AbstractVector methods)
# ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("nonuni", [False, True])
def test_1d_plain_centered(D, O, dtype, nonuni):
    n = 257
    for (d, a) in CENTERED:
        dx = nonuniform_dx(n, 0.01, dtype) if nonuni else 0.01
        for coeff in (1, 3.3, np.sin(np.arange(n) * 0.7)):
            A, B = make_pair("centered", d, a, dx, n, coeff, dtype=dtype)
            x = uniform_field(n + 2, dtype, seed=d * 10 + a)
            got = A * x
            want = O.apply_axis(B, x)
            assert got.shape == (n,)
            assert_close(got, want, dtype, f"centered ({d},{a}) nonuni={nonuni}")


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("nonuni", [False, True])
def test_1d_plain_upwind_mixed_sign(D, O, dtype, nonuni):
    n = 193
    for (d, a, off) in UPWIND:
        dx = nonuniform_dx(n, 0.02, dtype) if nonuni else 0.02
        csin = np.sin(6 * np.pi * np.arange(1, n + 1) / n)
        csin[::17] = 0.0                                            # exact zeros take the c >= 0 branch
        for coeff in (1.0, -1.0, 4.56, -4.56, csin, -csin):
            A, B = make_pair("upwind", d, a, dx, n, coeff, offside=off, dtype=dtype)
            x = uniform_field(n + 2, dtype, seed=d + a + off)
            assert_close(A * x, O.apply_axis(B, x), dtype, f"upwind ({d},{a}) off={off} nonuni={nonuni}")


# ------------------------------------------------------------------------------------------------------
# 1-D with boundary conditions:  (L*Q)*u   (BoundaryPaddedVector methods, convolutions.jl:367-730)
# ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", DTYPES)
def test_1d_ghost_all_bcs(D, O, dtype):
    n = 300
    h = 1.0 / (n + 1)
    for bspec in BCS:
        Qd, Qo = bc_pair(bspec, h, dtype)
        for (d, a) in [(2, 2), (2, 4), (1, 4), (2, 6), (4, 4)]:
            A, B = make_pair("centered", d, a, h, n, 1, dtype=dtype)
            u = uniform_field(n, dtype, seed=3)
            got = (A * Qd) * u
            want = O.apply_axis(B, u, Qo)
            assert_close(got, want, dtype, f"{bspec} centered ({d},{a})")
            # L * (Q*u) is the same application
            assert np.array_equal(A * (Qd * u), got)
        for (d, a, off) in [(1, 1, 0), (1, 2, 0), (2, 3, 0), (1, 3, 1)]:
            c = np.cos(np.arange(n) * 0.3)
            A, B = make_pair("upwind", d, a, h, n, c, offside=off, dtype=dtype)
            u = uniform_field(n, dtype, seed=4)
            assert_close((A * Qd) * u, O.apply_axis(B, u, Qo), dtype, f"{bspec} upwind ({d},{a},{off})")


@pytest.mark.parametrize("dtype", DTYPES)
def test_1d_ghost_vector_coefficients_bpv_indexing(D, O, dtype):
    """Spatially varying coefficients on a centered op with bpc > 0 through the BoundaryPaddedVector
    methods: the reference indexes coeff[i - bpc] there (convolutions.jl:384,:393,:428,:454)."""
    n = 120
    h = 0.05
    c = 1.0 + 0.5 * np.sin(np.arange(n) * 0.37)
    Qd, Qo = bc_pair(("robin", (1.0, 0.5, 0.25), (1.0, -0.5, 0.75), 1), h, dtype)
    for nonuni in (False, True):
        dx = nonuniform_dx(n, h, dtype) if nonuni else h
        for (d, a) in [(2, 4), (2, 6), (4, 4)]:
            A, B = make_pair("centered", d, a, dx, n, c, dtype=dtype)
            u = uniform_field(n, dtype, seed=5)
            assert_close((A * Qd) * u, O.apply_axis(B, u, Qo), dtype, f"bpv coeff ({d},{a}) nonuni={nonuni}")
            x = uniform_field(n + 2, dtype, seed=6)      # plain path: coeff[i], right rows coeff[1:bpc] (:109)
            assert_close(A * x, O.apply_axis(B, x), dtype, f"plain coeff ({d},{a}) nonuni={nonuni}")


def test_config1_heat_equation_1d(D, O):
    """BASELINE config 1 at reduced and full size: CenteredDifference(2,2,h,N)*Dirichlet0BC."""
    for n in (1000, 10 ** 6):
        h = 1.0 / (n + 1)
        A, B = make_pair("centered", 2, 2, h, n)
        Qd, Qo = bc_pair(("dirichlet0",), h, np.float64)
        u = uniform_field(n, np.float64, seed=n % 7)
        du = D.DeviceArray((n,), np.float64)
        ud = D.DeviceArray.from_host(u)
        D.mul_(du, A * Qd, ud)
        assert_close(du.to_host(), O.apply_axis(B, u, Qo), np.float64, f"C1 n={n}")
        plan = D.build_plans(A * Qd, (n,), (n,), np.float64)[0][0]
        assert plan.info[0] == "line"
        # single uniform operator: the 1-D kernel keeps the per-point kernel's arithmetic order -> identical bits
        dg = D.DeviceArray((n,), np.float64)
        D.mul_(dg, A * Qd, ud, flags=_flags(D, True))
        assert np.array_equal(dg.to_host(), du.to_host())


def test_line_kernel_selection_and_table_mode(D, O):
    """1-D plans: uniform + constant coefficient -> 'line' (weights in the constant bank); non-uniform grids,
    coefficient vectors, mixed-sign upwind and sums of operators -> 'line-table' (merged per-row weights);
    1-D PeriodicBC (l = u[end], r = u[1], bc_operators.jl:192) runs on the line kernel too (wrap-around read)."""
    n = 4099                                                           # odd length: scalar tail stores
    h = 1.0 / (n + 1)
    u = uniform_field(n, np.float64, seed=12)
    Qd, Qo = bc_pair(("robin", (1.0, 0.5, 0.25), (1.0, -0.5, 0.75), 1), h, np.float64)
    c = np.sin(6 * np.pi * np.arange(1, n + 1) / n)
    dx = nonuniform_dx(n, h, np.float64)
    A1, B1 = make_pair("centered", 2, 4, dx, n, 1)
    A2, B2 = make_pair("upwind", 1, 2, dx, n, c)
    A3, B3 = make_pair("centered", 1, 4, h, n, 1)
    for ops, want_kernel in [([(A3, B3)], "line"), ([(A1, B1)], "line-table"), ([(A2, B2)], "line-table"),
                             ([(A1, B1), (A2, B2), (A3, B3)], "line-table")]:
        A = ops[0][0]
        for o in ops[1:]:
            A = A + o[0]
        G = A * Qd
        assert D.build_plans(G, (n,), (n,), np.float64)[0][0].info[0] == want_kernel
        want = sum(O.apply_axis(b, u, Qo) for _, b in ops)
        assert_close(G * u, want, np.float64, f"line {want_kernel} {len(ops)} ops")
    for Ap, Bp in ((A3, B3), (A2, B2)):
        Gp = Ap * D.PeriodicBC(np.float64)
        assert D.build_plans(Gp, (n,), (n,), np.float64)[0][0].info[0].startswith("line")
        assert_close(Gp * u, O.apply_axis(Bp, u, O.PeriodicBC()), np.float64, "line periodic")
        gen = D.mul_alloc(Gp, D.DeviceArray.from_host(u), flags=D._lib.DEO_FLAG_FORCE_GENERIC).to_host()
        assert_close(Gp * u, gen, np.float64, "line periodic vs per-point kernel")
    # overwrite = false on the 1-D kernel
    du0 = uniform_field(n, np.float64, seed=13)
    dd = D.DeviceArray.from_host(du0)
    D.mul_(dd, A3 * Qd, D.DeviceArray.from_host(u), overwrite=False)
    assert_close(dd.to_host(), du0 + O.apply_axis(B3, u, Qo), np.float64, "line accumulate")


# ------------------------------------------------------------------------------------------------------
# golden vectors through the GPU: convert_by_multiplication with the CUDA mul!
# ------------------------------------------------------------------------------------------------------
def test_golden_operator_matrices_via_gpu(D, golden):
    for e in golden["operator_matrices"]:
        dx = e["dx"] if np.ndim(e["dx"]) == 0 else np.asarray(e["dx"])
        mk = D.CenteredDifference if e["kind"] == "centered" else D.UpwindDifference
        L = mk(e["d"], e["a"], dx, e["n"], e["coeff"])
        n = e["n"]
        eye = np.asfortranarray(np.eye(n + 2))                     # all unit vectors at once: an (n+2) x (n+2) "2-D array"
        got = L * eye                                              # axis-1 operator on every column
        want = np.asarray(e["matrix"])
        if "rows" in e:
            got = got[e["rows"][0]:e["rows"][1]]
        assert np.abs(got - want).max() <= 1e-8 * np.abs(want).max(), e["cite"]


def test_golden_ghost_matrices_via_gpu(D, golden):
    g = golden["ghost_operator_matrices"]
    dx, M = g["dx"], g["M"]
    for case in g["cases"]:
        Q = D.Neumann0BC(dx, 1) if case["bc"]["type"] == "neumann0" else D.RobinBC(case["bc"]["l"], case["bc"]["r"], dx, case["bc"]["order"])
        total = np.zeros((M, M))
        for t in case["terms"]:
            mk = D.CenteredDifference if t["kind"] == "centered" else D.UpwindDifference
            L = mk(t["d"], t["a"], dx, M, t["coeff"])
            if "scale" in t:
                L = t["scale"] * L
            affine = (L * Q) * np.zeros(M)
            total += np.stack([(L * Q) * e - affine for e in np.eye(M)], axis=1)
        total *= case.get("scale", 1.0)
        want = np.asarray(case["matrix"])
        assert np.abs(total - want).max() <= 1e-9 * np.abs(want).max() + 1e-12, case["name"]


# ------------------------------------------------------------------------------------------------------
# N-D single operator on pre-padded arrays (derivative_operator_functions.jl:27-69)
# ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", DTYPES)
def test_nd_prepadded_single_axis(D, O, dtype):
    for shape, axis in [((40, 23), 1), ((23, 40), 2), ((36, 17, 9), 1), ((17, 36, 9), 2), ((9, 17, 36), 3),
                        ((3, 2, 4, 2, 3, 40, 2), 6)]:
        n = shape[axis - 1]
        for kind, d, a, off in [("centered", 2, 2, 0), ("centered", 4, 4, 0), ("centered", 1, 4, 0), ("upwind", 1, 2, 0), ("upwind", 2, 3, 1)]:
            coeff = np.sin(np.arange(n) * 1.1) if kind == "upwind" else 1
            A, B = make_pair(kind, d, a, 0.1, n, coeff, axis=axis, offside=off, dtype=dtype)
            mshape = list(shape); mshape[axis - 1] += 2
            M = uniform_field(mshape, dtype, seed=len(shape) + axis)
            got = A * M
            assert got.shape == tuple(shape)
            assert_close(got, O.apply_axis(B, M), dtype, f"{shape} axis {axis} {kind} ({d},{a})")


def test_nd_prepadded_unnecessary_padding_is_stripped(D, O):
    """mul!(x_temp, A, M) with M padded on every dim: non-axis dims use 2:end-1 (derivative_operator_functions.jl:50-57)."""
    shape = (20, 14, 11)
    M = uniform_field([s + 2 for s in shape], np.float64, seed=9)
    for axis in (1, 2, 3):
        A, B = make_pair("centered", 2, 4, 0.1, shape[axis - 1], axis=axis)
        du = np.zeros(shape, order="F")
        D.mul_(du, A, M)
        want = O.apply_axis(B, M, out=np.zeros(shape, order="F"))
        assert_close(du, want, np.float64, f"strip axis {axis}")


def test_composite_prepadded_2d_3d(D, O):
    """mul!(x_temp, Dxx+Dyy[+Dzz], M) on a dense pre-padded array (derivative_operator_functions.jl:203,:466),
    intended semantics: each op reads its own axis' ghost layer, 2:end-1 on the others."""
    for shape in [(31, 27), (19, 23, 17)]:
        nd = len(shape)
        M = uniform_field([s + 2 for s in shape], np.float64, seed=11)
        pairs = [make_pair("centered", 2, 4 if ax < 3 else 2, 0.1 * ax, shape[ax - 1], axis=ax) for ax in range(1, nd + 1)]
        A = pairs[0][0]
        for p in pairs[1:]:
            A = A + p[0]
        du = np.zeros(shape, order="F")
        D.mul_(du, A, M)
        want = O.apply_sum([p[1] for p in pairs], M, None)
        assert_close(du, want, np.float64, f"composite prepadded {shape}")


# ------------------------------------------------------------------------------------------------------
# N-D with boundary conditions: (sum L)*Q*u   (ghost_derivative_operator.jl:11-24, composite_operators.jl:64-65)
# ------------------------------------------------------------------------------------------------------
def _laplacian_pair(shape, a, h, dtype, coeff=1):
    pairs = [make_pair("centered", 2, a, h[ax - 1], shape[ax - 1], coeff, axis=ax, dtype=dtype) for ax in range(1, len(shape) + 1)]
    A = pairs[0][0]
    for p in pairs[1:]:
        A = A + p[0]
    return A, [p[1] for p in pairs]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("generic", [False, True])
def test_config2_2d_laplacian_robin(D, O, dtype, generic):
    """BASELINE config 2 at reduced size: Dxx+Dyy (2,4) with RobinBC, uniform grid."""
    for shape in [(256, 192), (130, 70), (64, 1030)]:
        h = (1.0 / (shape[0] + 1), 1.0 / (shape[1] + 1))
        A, Bs = _laplacian_pair(shape, 4, h, dtype)
        Q = D.compose(*D.RobinBC((1.0, 0.5, 0.25), (1.0, -0.5, 0.75), h, 1, shape, dtype=dtype))
        bcs = {ax: O.RobinBC((1.0, 0.5, 0.25), (1.0, -0.5, 0.75), h[ax - 1], 1, dtype) for ax in (1, 2)}
        u = uniform_field(shape, dtype, seed=21)
        got = D.mul_alloc(A * Q, u, flags=_flags(D, generic))
        assert_close(got, O.apply_sum(Bs, u, bcs), dtype, f"C2 {shape}")


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("generic", [False, True])
def test_config3_3d_laplacian_neumann(D, O, dtype, generic):
    """BASELINE config 3 at reduced size: Dxx+Dyy+Dzz (2,6) with MultiDimBC Neumann (zero and non-zero)."""
    for shape in [(64, 48, 40), (70, 33, 29), (128, 16, 20)]:
        h = tuple(1.0 / (s + 1) for s in shape)
        A, Bs = _laplacian_pair(shape, 6, h, dtype)
        for alpha in [(0.0, 0.0), (0.3, -0.2)]:
            if alpha == (0.0, 0.0):
                Q = D.compose(*D.Neumann0BC(dtype, h, 1, shape))
            else:
                Q = D.compose(*D.NeumannBC(alpha, h, 1, shape, dtype=dtype))
            bcs = {ax: O.NeumannBC(alpha, h[ax - 1], 1, dtype) for ax in (1, 2, 3)}
            u = uniform_field(shape, dtype, seed=31)
            got = D.mul_alloc(A * Q, u, flags=_flags(D, generic))
            assert_close(got, O.apply_sum(Bs, u, bcs), dtype, f"C3 {shape} alpha={alpha}")


@pytest.mark.parametrize("generic", [False, True])
def test_3d_laplacian_dirichlet_and_orders(D, O, generic):
    """test/DerivativeOperators/3D_laplacian.jl:10-20 operator (2,4) + Dirichlet0; plus (2,2), (2,8)."""
    shape = (51, 51, 51)
    h = (0.2, 0.2, 0.2)
    for a in (2, 4, 8):
        A, Bs = _laplacian_pair(shape, a, h, np.float64)
        Q = D.compose(*D.Dirichlet0BC(np.float64, shape))
        bcs = {ax: O.Dirichlet0BC(np.float64) for ax in (1, 2, 3)}
        u = uniform_field(shape, np.float64, seed=a)
        got = D.mul_alloc(A * Q, u, flags=_flags(D, generic))
        assert_close(got, O.apply_sum(Bs, u, bcs), np.float64, f"3D laplacian a={a}")


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("generic", [False, True])
def test_config4_nonuniform_centered_plus_upwind(D, O, dtype, generic):
    """BASELINE config 4 at reduced size: non-uniform grid, centered (1,4) + (2,4) and upwind (1,2) with a
    variable-sign coefficient field (exact zeros included), RobinBC built from the same dx vectors."""
    shape = (48, 40, 36)
    hs = [1.0 / (s + 1) for s in shape]
    dxs = [nonuniform_dx(s, h, dtype) for s, h in zip(shape, hs)]
    cs = []
    for s in shape:
        c = np.sin(6 * np.pi * np.arange(1, s + 1) / s)
        c[::8] = 0.0
        cs.append(c)
    fam = {
        "lap": [make_pair("centered", 2, 4, dxs[ax - 1], shape[ax - 1], axis=ax, dtype=dtype) for ax in (1, 2, 3)],
        "grad": [make_pair("centered", 1, 4, dxs[ax - 1], shape[ax - 1], axis=ax, dtype=dtype) for ax in (1, 2, 3)],
        "adv": [make_pair("upwind", 1, 2, dxs[ax - 1], shape[ax - 1], cs[ax - 1], axis=ax, dtype=dtype) for ax in (1, 2, 3)],
    }
    Q = D.compose(*D.RobinBC((1.0, 0.5, 0.25), (1.0, -0.5, 0.75), dxs, 1, shape, dtype=dtype))
    bcs = {ax: O.RobinBC((1.0, 0.5, 0.25), (1.0, -0.5, 0.75), dxs[ax - 1], 1, dtype) for ax in (1, 2, 3)}
    u = uniform_field(shape, dtype, seed=41)
    for name, pairs in list(fam.items()) + [("all9", fam["lap"] + fam["grad"] + fam["adv"])]:
        A = pairs[0][0]
        for p in pairs[1:]:
            A = A + p[0]
        got = D.mul_alloc(A * Q, u, flags=_flags(D, generic))
        assert_close(got, O.apply_sum([p[1] for p in pairs], u, bcs), dtype, f"C4 {name}")
        # the tiled kernel takes non-uniform / upwind / multi-operator sums through its merged per-row tables
        kern = D.build_plans(A * Q, shape, shape, dtype, flags=_flags(D, generic))[0][0].info[0]
        assert kern == ("generic" if generic else "star-table"), kern


def test_nd_single_directional_bc(D, O):
    """Dxx * Qx * u with a MultiDimDirectionalBC (only that axis padded), every axis."""
    shape = (30, 26, 22)
    u = uniform_field(shape, np.float64, seed=51)
    for ax in (1, 2, 3):
        A, B = make_pair("centered", 2, 4, 0.1, shape[ax - 1], axis=ax)
        Qd = D.RobinBC[ax]((1.0, 2.0, 3.0), (0.0, -1.0, 2.0), 0.1, 4, shape)
        Qo = O.RobinBC((1.0, 2.0, 3.0), (0.0, -1.0, 2.0), 0.1, 4)
        assert_close((A * Qd) * u, O.apply_axis(B, u, Qo), np.float64, f"directional axis {ax}")
    # an atomic BC applied to an array extends dimension 1 (multi_dim_bc_operators.jl:210)
    A, B = make_pair("centered", 2, 2, 0.1, shape[0], axis=1)
    Qd, Qo = bc_pair(("neumann", (0.3, -0.2), 1), 0.1, np.float64)
    assert_close((A * Qd) * u, O.apply_axis(B, u, Qo), np.float64, "atomic BC on array")


def test_per_face_bc_tables(D, O):
    """MultiDimBC{dim}(array of BCs): one atomic BC per boundary pencil, different K per pencil."""
    shape = (20, 14, 9)
    rng = np.random.default_rng(61)
    u = uniform_field(shape, np.float64, seed=62)
    for ax in (1, 2, 3):
        face = tuple(s for i, s in enumerate(shape) if i != ax - 1)
        arr = np.empty(face, dtype=object)
        a_l = np.zeros(face + (3,)); a_r = np.zeros(face + (3,)); b_l = np.zeros(face); b_r = np.zeros(face)
        for idx in np.ndindex(*face):
            order = int(rng.integers(1, 4))
            l = tuple(rng.uniform(0.5, 2.0, 3)); r = tuple(rng.uniform(0.5, 2.0, 3))
            q = D.RobinBC(l, r, 0.1, order)
            arr[idx] = q
            a_l[idx][:order] = q.a_l; a_r[idx][3 - order:] = q.a_r; b_l[idx] = q.b_l; b_r[idx] = q.b_r
        Qd = D.MultiDimBC[ax](arr)
        nface = int(np.prod(face))
        Qo = O.BC(a_l.reshape((nface, 3), order="F"), b_l.reshape(-1, order="F"),
                  a_r.reshape((nface, 3), order="F"), b_r.reshape(-1, order="F"), np.float64)
        A, B = make_pair("centered", 2, 4, 0.1, shape[ax - 1], axis=ax)
        assert_close((A * Qd) * u, O.apply_axis(B, u, Qo), np.float64, f"per-face axis {ax}")


def _kernel_of(D, G, out_shape, in_shape, dtype):
    return D.apply._get_plans(G, out_shape, in_shape, dtype, False, 0)[0][0].info[0]


@pytest.mark.parametrize("dtype", DTYPES)
def test_tiled_kernel_takes_prepadded_perface_and_periodic_inputs(D, O, dtype):
    """The shapes that used to fall back to the per-point kernel on the tiled kernel, several tiles wide
    and tall: pre-padded dense input (derivative_operator_functions.jl:203,:466), per-pencil MultiDimBC arrays
    (multi_dim_bc_operators.jl:54-57), N-D PeriodicBC (:221-228)."""
    rng = np.random.default_rng(3)
    for shape, a in [((200, 70, 37), 4), ((136, 44, 30), 6), ((264, 90), 4)]:
        nd = len(shape)
        h = tuple(1.0 / (s + 1) for s in shape)
        # --- pre-padded composite: each operator reads its own axis' ghost layer of M (the contiguous axis included: its
        # rows are one element behind the rows of du, so the kernel shifts its tile origins to keep the TMA boxes aligned, or
        # copies element-wise when the padded row pitch is not a multiple of 16 bytes)
        pairs = [make_pair("centered", 2, a, h[ax - 1], shape[ax - 1], axis=ax, dtype=dtype) for ax in range(1, nd + 1)]
        M = uniform_field([s + 2 for s in shape], dtype, seed=21)
        A = pairs[0][0]
        for pr in pairs[1:]:
            A = A + pr[0]
        du = np.zeros(shape, dtype=dtype, order="F")
        D.mul_(du, A, M)
        assert _kernel_of(D, A, shape, M.shape, dtype).startswith("star"), "pre-padded input must run tiled"
        assert_close(du, O.apply_sum([pr[1] for pr in pairs], M, None), dtype, f"pre-padded {shape}")
        # --- N-D periodic
        u = uniform_field(shape, dtype, seed=22)
        Qp = D.compose(*D.PeriodicBC(dtype, shape))
        assert _kernel_of(D, A * Qp, shape, shape, dtype).startswith("star"), "N-D PeriodicBC must run tiled"
        assert_close((A * Qp) * u, O.apply_sum([pr[1] for pr in pairs], u, {ax: O.PeriodicBC(dtype) for ax in range(1, nd + 1)}), dtype,
                     f"periodic {shape}")
    # --- per-pencil BC tables on every axis of a 3-D array (different Robin data and stencil length per pencil)
    shape, a = (136, 44, 30), 4
    h = tuple(1.0 / (s + 1) for s in shape)
    u = uniform_field(shape, dtype, seed=23)
    Qd, Qo = [], {}
    for ax in (1, 2, 3):
        face = tuple(s for i, s in enumerate(shape) if i != ax - 1)
        arr = np.empty(face, dtype=object)
        a_l = np.zeros(face + (3,), dtype=dtype); a_r = np.zeros(face + (3,), dtype=dtype)
        b_l = np.zeros(face, dtype=dtype); b_r = np.zeros(face, dtype=dtype)
        for idx in np.ndindex(*face):
            order = int(rng.integers(1, 4))
            q = D.RobinBC(tuple(rng.uniform(0.5, 2.0, 3)), tuple(rng.uniform(0.5, 2.0, 3)), h[ax - 1], order, dtype=dtype)
            arr[idx] = q
            a_l[idx][:order] = q.a_l; a_r[idx][3 - order:] = q.a_r; b_l[idx] = q.b_l; b_r[idx] = q.b_r
        Qd.append(D.MultiDimBC[ax](arr))
        nface = int(np.prod(face))
        Qo[ax] = O.BC(a_l.reshape((nface, 3), order="F"), b_l.reshape(-1, order="F"), a_r.reshape((nface, 3), order="F"),
                      b_r.reshape(-1, order="F"), dtype)
    pairs = [make_pair("centered", 2, a, h[ax - 1], shape[ax - 1], axis=ax, dtype=dtype) for ax in (1, 2, 3)]
    A = pairs[0][0] + pairs[1][0] + pairs[2][0]
    G = A * D.compose(*Qd)
    assert _kernel_of(D, G, shape, shape, dtype).startswith("star"), "per-pencil BC arrays must run tiled"
    assert_close(G * u, O.apply_sum([pr[1] for pr in pairs], u, Qo), dtype, "per-pencil BC tables, 3 axes")


def _per_pencil_bcs(D, O, shape, h, dtype, rng):
    """One RobinBC per boundary pencil on every axis (different data and stencil length per pencil): product MultiDimBC list and
    the oracle's per-face tables."""
    Qd, Qo = [], {}
    for ax in range(1, len(shape) + 1):
        face = tuple(s for i, s in enumerate(shape) if i != ax - 1)
        arr = np.empty(face, dtype=object)
        a_l = np.zeros(face + (3,), dtype=dtype); a_r = np.zeros(face + (3,), dtype=dtype)
        b_l = np.zeros(face, dtype=dtype); b_r = np.zeros(face, dtype=dtype)
        for idx in np.ndindex(*face):
            order = int(rng.integers(1, 4))
            q = D.RobinBC(tuple(rng.uniform(0.5, 2.0, 3)), tuple(rng.uniform(0.5, 2.0, 3)), h[ax - 1], order, dtype=dtype)
            arr[idx] = q
            a_l[idx][:order] = q.a_l; a_r[idx][3 - order:] = q.a_r; b_l[idx] = q.b_l; b_r[idx] = q.b_r
        Qd.append(D.MultiDimBC[ax](arr))
        nface = int(np.prod(face))
        Qo[ax] = O.BC(a_l.reshape((nface, 3), order="F"), b_l.reshape(-1, order="F"), a_r.reshape((nface, 3), order="F"),
                      b_r.reshape(-1, order="F"), dtype)
    return Qd, Qo


@pytest.mark.parametrize("dtype", DTYPES)
def test_per_pencil_bc_tables_on_shifted_tiles(D, O, dtype):
    """Per-pencil MultiDimBC arrays (multi_dim_bc_operators.jl:54-57) on extents just above tile multiples: the face index
    of a pencil must not depend on where the tile origins are; also odd row lengths (cp.async loader) and 2-D strips."""
    rng = np.random.default_rng(5)
    vec = 16 // np.dtype(dtype).itemsize
    tx, ty = 32 * vec, 32
    for shape, a in [((tx + 2, ty + 1, 38), 4), ((2 * tx + 1, 2 * ty + 2, 40), 2), ((tx - 2 * vec, 70, 37), 6), ((32 * vec * 32 + 2, 41), 4)]:
        h = tuple(1.0 / (s + 1) for s in shape)
        u = uniform_field(shape, dtype, seed=25)
        Qd, Qo = _per_pencil_bcs(D, O, shape, h, dtype, rng)
        pairs = [make_pair("centered", 2, a, h[ax - 1], shape[ax - 1], axis=ax, dtype=dtype) for ax in range(1, len(shape) + 1)]
        A = pairs[0][0]
        for pr in pairs[1:]:
            A = A + pr[0]
        G = A * D.compose(*Qd)
        assert _kernel_of(D, G, shape, shape, dtype).startswith("star"), f"{shape}: per-pencil BC arrays must run tiled"
        assert_close(G * u, O.apply_sum([pr[1] for pr in pairs], u, Qo), dtype, f"per-pencil BC tables {shape}")


@pytest.mark.parametrize("dtype", DTYPES)
def test_prepadded_contiguous_axis_runs_on_shifted_tiles(D, O, dtype):
    """mul!(x_temp, A, M) with M padded along the contiguous axis (derivative_operator_functions.jl:27-69,:203,:466; the shape
    of every Gradient / Divergence / Curl input): tile origins shifted so that the boxes of M start on 16-byte boundaries, du
    written element-wise.  Row lengths that are a multiple of the tile width (one more tile column), narrow first / last
    tiles, 2-D strips, non-uniform grids (TABLE), overwrite=false."""
    vec = 16 // np.dtype(dtype).itemsize
    tx = 32 * vec
    shapes = [((2 * tx + (0 if vec == 2 else 2), 40, 41), 4), ((2 * tx - 2, 37, 39), 2), ((3 * tx + 6, 45, 38), 6), ((tx + 2 * vec - 2, 70, 37), 4), ((40 * tx - 2, 43), 4)]
    for shape, a in shapes:
        assert ((shape[0] + 2) * np.dtype(dtype).itemsize) % 16 == 0       # tensor map possible: this is the shifted-tile path
        nd = len(shape)
        h = tuple(1.0 / (s + 1) for s in shape)
        M = uniform_field([s + 2 for s in shape], dtype, seed=23)
        for nonuni in (False, True):
            pairs = [make_pair("centered", 2, a, nonuniform_dx(shape[ax - 1], h[ax - 1], dtype) if nonuni else h[ax - 1], shape[ax - 1], axis=ax, dtype=dtype)
                     for ax in range(1, nd + 1)]
            A = pairs[0][0]
            for pr in pairs[1:]:
                A = A + pr[0]
            kern = _kernel_of(D, A, shape, M.shape, dtype)
            assert kern == ("star-table" if nonuni else "star"), f"{shape} nonuni={nonuni}: {kern}"
            want = O.apply_sum([pr[1] for pr in pairs], M, None)
            du = np.zeros(shape, dtype=dtype, order="F")
            D.mul_(du, A, M)
            assert_close(du, want, dtype, f"pre-padded, shifted tiles {shape} nonuni={nonuni}")
            # one operator along the contiguous axis only (a Gradient component)
            dx1 = np.zeros(shape, dtype=dtype, order="F")
            D.mul_(dx1, pairs[0][0], M)
            assert_close(dx1, O.apply_axis(pairs[0][1], M, out=np.zeros(shape, dtype=dtype, order="F")), dtype, f"pre-padded x operator {shape} nonuni={nonuni}")
        base = uniform_field(shape, dtype, seed=24)
        acc = D.DeviceArray.from_host(base)
        D.mul_(acc, A, D.DeviceArray.from_host(M), overwrite=False)
        assert np.abs(acc.to_host().astype(np.float64) - (base.astype(np.float64) + want)).max() <= TOL[np.dtype(dtype)] * np.abs(want).max()


@pytest.mark.parametrize("dtype", DTYPES)
def test_extents_just_above_a_tile_multiple_run_tiled(D, O, dtype):
    """Extents of the form k * tile + 1, + 2 (the 2^k + 1 grids): the last tile along x / y alone is too narrow to hold its
    face's boundary stencil; the kernel shifts its tile origins instead of falling back to the per-point kernel.  Even and
    odd row lengths (TMA / cp.async loader), Robin and Dirichlet faces, non-uniform grids, overwrite=false."""
    vec = 16 // np.dtype(dtype).itemsize
    tx, ty = 32 * vec, 32
    cases = [((2 * tx + 2, 2 * ty + 1, 41), 4), ((tx + 1, ty + 2, 39), 4), ((3 * tx + 2, 3 * ty + 3, 38), 6), ((2 * tx + 1, 70, 40), 2),
             ((32 * vec * 32 + 2, 45), 4), ((32 * vec * 32 + 1, 41), 6)]
    for shape, a in cases:
        nd = len(shape)
        h = tuple(1.0 / (s + 1) for s in shape)
        u = uniform_field(shape, dtype, seed=31)
        for bc, nonuni in (("robin", False), ("dirichlet0", False), ("robin", True)):
            dxs = [nonuniform_dx(shape[ax], h[ax], dtype) if nonuni else h[ax] for ax in range(nd)]
            pairs = [make_pair("centered", 2, a, dxs[ax - 1], shape[ax - 1], axis=ax, dtype=dtype) for ax in range(1, nd + 1)]
            A = pairs[0][0]
            for pr in pairs[1:]:
                A = A + pr[0]
            if bc == "dirichlet0":
                Q = D.compose(*D.Dirichlet0BC(dtype, shape))
                bcs = {ax + 1: O.Dirichlet0BC(dtype) for ax in range(nd)}
            else:
                Q = D.compose(*D.RobinBC((1.0, 0.5, 0.25), (1.0, -0.5, 0.75), dxs if nonuni else h, 1, shape, dtype=dtype))
                bcs = {ax + 1: O.RobinBC((1.0, 0.5, 0.25), (1.0, -0.5, 0.75), dxs[ax], 1, dtype) for ax in range(nd)}
            G = A * Q
            assert _kernel_of(D, G, shape, shape, dtype).startswith("star"), f"{shape} {bc} nonuni={nonuni} must run tiled"
            want = O.apply_sum([pr[1] for pr in pairs], u, bcs)
            assert_close(G * u, want, dtype, f"shifted tiles {shape} {bc} nonuni={nonuni}")
        acc = D.DeviceArray.from_host(u)
        D.mul_(acc, G, D.DeviceArray.from_host(u), overwrite=False)
        assert np.abs(acc.to_host().astype(np.float64) - (u.astype(np.float64) + want)).max() <= TOL[np.dtype(dtype)] * np.abs(want).max()


@pytest.mark.parametrize("dtype", DTYPES)
def test_seeded_random_1d_operators(D, O, dtype):
    """Seeded sweep of the 1-D kernels: random lengths (a few points above the smallest legal grid up to several thousand),
    every centered / upwind family of the lists above, uniform and non-uniform grids, constant / mixed-sign coefficient
    vectors, every BC family (PeriodicBC included) or a plain padded vector, sums of operators, overwrite=false."""
    rng = np.random.default_rng(777 + np.dtype(dtype).itemsize)
    for case in range(60):
        n = int(rng.choice([rng.integers(24, 64), rng.integers(64, 700), rng.integers(700, 6000)]))
        h = 1.0 / (n + 1)
        nonuni = bool(rng.integers(0, 2))
        dx = nonuniform_dx(n, h, dtype) if nonuni else h
        nops = int(rng.integers(1, 4))
        pairs = []
        for _ in range(nops):
            coeff = [1.0, -2.5, np.sin(7 * np.pi * np.arange(1, n + 1) / n)][int(rng.integers(0, 3))]
            if rng.integers(0, 2):
                d, a = CENTERED[int(rng.integers(0, len(CENTERED) - 2))]            # reach <= 4 and the (4,10) / (8,8) giants separately below
                pairs.append(make_pair("centered", d, a, dx, n, coeff, dtype=dtype))
            else:
                d, a, off = UPWIND[int(rng.integers(0, len(UPWIND)))]
                pairs.append(make_pair("upwind", d, a, dx, n, coeff, offside=off, dtype=dtype))
        A = pairs[0][0]
        for pr in pairs[1:]:
            A = A + pr[0]
        what = f"case {case}: n={n} nonuni={nonuni} ops={[type(p[0]).__name__ for p in pairs]}"
        if rng.integers(0, 4) == 0:
            x = uniform_field(n + 2, dtype, seed=300 + case)                           # plain padded vector
            want = O.apply_axis(pairs[0][1], x)
            for pr in pairs[1:]:
                want = want + O.apply_axis(pr[1], x)
            assert_close(A * x, want, dtype, what + " padded")
            continue
        bspec = BCS[int(rng.integers(0, len(BCS)))]
        Qd, Qo = bc_pair(bspec, dx if (nonuni and bspec[0] in ("robin", "neumann", "general")) else h, dtype)
        u = uniform_field(n, dtype, seed=300 + case)
        want = O.apply_axis(pairs[0][1], u, Qo)
        for pr in pairs[1:]:
            want = want + O.apply_axis(pr[1], u, Qo)
        assert_close((A * Qd) * u, want, dtype, what + f" {bspec[0]}")
        if case % 4 == 0:
            base = uniform_field(n, dtype, seed=400 + case)
            acc = D.DeviceArray.from_host(base)
            D.mul_(acc, A * Qd, D.DeviceArray.from_host(u), overwrite=False)
            err = np.abs(acc.to_host().astype(np.float64) - (base.astype(np.float64) + want)).max() / max(np.abs(want).max(), 1e-300)
            assert err <= TOL[np.dtype(dtype)], f"{what} overwrite=false: {err:.3e}"


@pytest.mark.parametrize("dtype", DTYPES)
def test_seeded_random_shapes_on_the_tiled_kernel(D, O, dtype):
    """A seeded sweep over what the tiling has to get right at once: extents on both sides of tile multiples (shifted
    origins along x and / or y, narrow first and last tiles), even and odd row lengths (TMA / cp.async loader), 2-D strips
    and 3-D tiles, radius 1..3, uniform and non-uniform grids, an upwind term with a mixed-sign coefficient vector, affine
    faces or a pre-padded input, overwrite=false -- every result against the oracle."""
    rng = np.random.default_rng(20260 + np.dtype(dtype).itemsize)
    vec = 16 // np.dtype(dtype).itemsize
    tx, ty = 32 * vec, 32
    ran_tiled, not_tiled = 0, []
    for case in range(28):
        nd = 3 if case % 4 else 2
        near = lambda t, kmax: int(rng.integers(1, kmax + 1)) * t + int(rng.integers(-3, 4))
        nx = max(near(tx, 3), 40) if nd == 3 else max(near(32 * vec * 32, 2), 40)
        shape = (nx, max(near(ty, 3), 40), int(rng.integers(37, 60))) if nd == 3 else (nx, int(rng.integers(37, 80)))
        a = int(rng.choice([2, 4, 6]))
        nonuni = bool(rng.integers(0, 2))
        upwind = bool(rng.integers(0, 2))
        padded = bool(rng.integers(0, 3) == 0)
        h = [1.0 / (s + 1) for s in shape]
        dxs = [nonuniform_dx(shape[ax], h[ax], dtype) if nonuni else h[ax] for ax in range(nd)]
        pairs = [make_pair("centered", 2, a, dxs[ax - 1], shape[ax - 1], axis=ax, dtype=dtype) for ax in range(1, nd + 1)]
        if upwind:
            c = np.sin(5 * np.pi * np.arange(1, shape[0] + 1) / shape[0]); c[::7] = 0.0
            pairs.append(make_pair("upwind", 1, 2, dxs[0], shape[0], c, axis=1, dtype=dtype))
        A = pairs[0][0]
        for pr in pairs[1:]:
            A = A + pr[0]
        what = f"case {case}: {shape} a={a} nonuni={nonuni} upwind={upwind} padded={padded}"
        if padded:
            M = uniform_field([s + 2 for s in shape], dtype, seed=100 + case)
            want = O.apply_sum([pr[1] for pr in pairs], M, None)
            G, u_host, in_shape = A, M, M.shape
        else:
            kind = int(rng.integers(0, 3))
            if kind == 0:
                Q = D.compose(*D.Dirichlet0BC(dtype, shape)); bcs = {ax + 1: O.Dirichlet0BC(dtype) for ax in range(nd)}
            elif kind == 1:
                Q = D.compose(*D.RobinBC((1.0, 0.5, 0.25), (1.0, -0.5, 0.75), dxs if nonuni else tuple(h), 1, shape, dtype=dtype))
                bcs = {ax + 1: O.RobinBC((1.0, 0.5, 0.25), (1.0, -0.5, 0.75), dxs[ax], 1, dtype) for ax in range(nd)}
            else:
                Q = D.compose(*D.RobinBC((1.0, 6.0, 10.0), (1.0, 6.0, 10.0), dxs if nonuni else tuple(h), 3, shape, dtype=dtype))
                bcs = {ax + 1: O.RobinBC((1.0, 6.0, 10.0), (1.0, 6.0, 10.0), dxs[ax], 3, dtype) for ax in range(nd)}
            u_host = uniform_field(shape, dtype, seed=100 + case)
            want = O.apply_sum([pr[1] for pr in pairs], u_host, bcs)
            G, in_shape = A * Q, shape
        if _kernel_of(D, G, shape, in_shape, dtype).startswith("star"):
            ran_tiled += 1
        else:
            not_tiled.append(what)
        du = np.zeros(shape, dtype=dtype, order="F")
        D.mul_(du, G, u_host)
        assert_close(du, want, dtype, what)
        if case % 3 == 0:
            base = uniform_field(shape, dtype, seed=200 + case)
            acc = D.DeviceArray.from_host(base)
            D.mul_(acc, G, D.DeviceArray.from_host(u_host), overwrite=False)
            err = np.abs(acc.to_host().astype(np.float64) - (base.astype(np.float64) + want)).max() / np.abs(want).max()
            assert err <= TOL[np.dtype(dtype)], f"{what} overwrite=false: {err:.3e}"
    assert ran_tiled >= 24, f"only {ran_tiled} of 28 cases ran on the tiled kernel; per-point kernel: {not_tiled}"


@pytest.mark.parametrize("dtype", DTYPES)
def test_tiled_kernel_takes_any_row_length(D, O, dtype):
    """Row lengths that are not a multiple of 16 bytes (no tensor map possible): cp.async element copies into the same
    shared-memory layout, element-wise stores of du.  Includes the reference's own 51^3 example
    (test/DerivativeOperators/3D_laplacian.jl:3-20: CenteredDifference(2,2) on 51^3 with Dirichlet0 faces)."""
    for shape, a, bc in [((51, 51, 51), 2, "dirichlet0"), ((67, 45, 33), 4, "robin"), ((131, 70, 29), 6, "robin"), ((333, 41), 4, "robin"),
                         ((2051, 37), 2, "robin")]:
        nd = len(shape)
        h = tuple(1.0 / (s + 1) for s in shape)
        A, Bs = _laplacian_pair(shape, a, h, dtype)
        if bc == "dirichlet0":
            Q = D.compose(*D.Dirichlet0BC(dtype, shape))
            bcs = {ax + 1: O.Dirichlet0BC(dtype) for ax in range(nd)}
        else:
            Q = D.compose(*D.RobinBC((1.0, 0.5, 0.25), (1.0, -0.5, 0.75), h, 1, shape, dtype=dtype))
            bcs = {ax + 1: O.RobinBC((1.0, 0.5, 0.25), (1.0, -0.5, 0.75), h[ax], 1, dtype) for ax in range(nd)}
        u = uniform_field(shape, dtype, seed=41)
        G = A * Q
        assert _kernel_of(D, G, shape, shape, dtype).startswith("star"), f"{shape} must run tiled"
        want = O.apply_sum(Bs, u, bcs)
        assert_close(G * u, want, dtype, f"odd row length {shape}")
        ud = D.DeviceArray.from_host(u)
        acc = D.DeviceArray.from_host(u)
        D.mul_(acc, G, ud, overwrite=False)                      # element-wise read-modify-write of du
        assert np.abs(acc.to_host().astype(np.float64) - (u.astype(np.float64) + want)).max() <= TOL[np.dtype(dtype)] * np.abs(want).max()


def test_periodic_nd_quirk(D, O):
    """N-D PeriodicBC ghosts are lower=u[1,...], upper=u[end,...] (multi_dim_bc_operators.jl:221-228),
    the reverse of the 1-D rule (bc_operators.jl:192); reproduced as is."""
    shape = (16, 12)
    u = uniform_field(shape, np.float64, seed=71)
    A1, B1 = make_pair("centered", 2, 2, 0.1, 16, axis=1)
    A2, B2 = make_pair("centered", 2, 2, 0.1, 12, axis=2)
    Q = D.compose(*D.PeriodicBC(np.float64, shape))
    got = ((A1 + A2) * Q) * u
    want = O.apply_sum([B1, B2], u, {1: O.PeriodicBC(), 2: O.PeriodicBC()})
    assert_close(got, want, np.float64, "periodic N-D")
    pad = (Q * u).to_array()
    assert np.array_equal(pad[0, 1:-1], u[0, :]) and np.array_equal(pad[-1, 1:-1], u[-1, :])


# ------------------------------------------------------------------------------------------------------
# accumulate, coefficient updates, host-buffer path, errors, edge sizes
# ------------------------------------------------------------------------------------------------------
def test_overwrite_false_accumulates(D, O):
    shape = (33, 21, 18)
    A, B = make_pair("centered", 2, 4, 0.1, shape[1], axis=2)
    M = uniform_field((33, 23, 18), np.float64, seed=81)
    old = uniform_field(shape, np.float64, seed=82)
    du = old.copy(order="F")
    D.mul_(du, A, M, overwrite=False)
    want = O.apply_axis(B, M, out=old.copy(order="F"), overwrite=False)
    assert_close(du, want, np.float64, "overwrite=false")


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape,a", [((96, 70, 41), 4), ((200, 37, 30), 6), ((300, 75), 4), ((130, 65, 40), 4), ((257, 66, 38), 2), ((4098, 45), 4)])
def test_tiled_accumulate_and_fused_axpy(D, O, dtype, shape, a):
    """overwrite = false (convolutions.jl:17-22) and the fused explicit-stepper update u + dt*(A u)
    (3D_laplacian.jl:20-24) on the tiled kernels: every face, every march-axis edge row."""
    nd = len(shape)
    h = tuple(1.0 / (s + 1) for s in shape)
    A, Bs = _laplacian_pair(shape, a, h, dtype)
    Q = D.compose(*D.RobinBC((1.0, 0.5, 0.25), (1.0, -0.5, 0.75), h, 1, shape, dtype=dtype))
    bcs = {ax + 1: O.RobinBC((1.0, 0.5, 0.25), (1.0, -0.5, 0.75), h[ax], 1, dtype) for ax in range(nd)}
    u = uniform_field(shape, dtype, seed=5)
    old = uniform_field(shape, dtype, seed=6)
    want = O.apply_sum(Bs, u, bcs).astype(np.float64)
    G = A * Q
    ud = D.DeviceArray.from_host(u)
    # accumulate
    dud = D.DeviceArray.from_host(old)
    D.mul_(dud, G, ud, overwrite=False)
    plan = D.apply._get_plans(G, shape, shape, dtype, True, 0)[0][0]
    assert plan.info[0].startswith("star"), plan.info
    scale = np.abs(want).max()
    assert np.abs(dud.to_host() - (old.astype(np.float64) + want)).max() <= TOL[np.dtype(dtype)] * scale
    # fused AXPY
    dt = 1e-6
    out = D.DeviceArray(shape, dtype)
    D.step_(out, G, ud, dt)
    ref = u.astype(np.float64) + dt * want
    assert np.abs(out.to_host() - ref).max() <= TOL[np.dtype(dtype)] * max(np.abs(ref).max(), dt * scale)
    # ... and through a plan that does not fuse (per-point kernel + update kernel)
    out2 = D.DeviceArray(shape, dtype)
    D.step_(out2, G, ud, dt, flags=D._lib.DEO_FLAG_FORCE_GENERIC)
    assert np.abs(out2.to_host() - ref).max() <= TOL[np.dtype(dtype)] * max(np.abs(ref).max(), dt * scale)


def test_update_coefficients_and_scalar_scaling(D, O):
    n = 200
    u = uniform_field(n, np.float64, seed=91)
    Qd, Qo = bc_pair(("robin", (1.0, 0.5, 0.25), (1.0, -0.5, 0.75), 1), 0.1, np.float64)
    c0 = np.sin(np.arange(n) * 0.2)
    A, B = make_pair("upwind", 1, 2, 0.1, n, c0)
    G = A * Qd
    first = G * u
    assert_close(first, O.apply_axis(B, u, Qo), np.float64, "before update")
    c1 = -np.cos(np.arange(n) * 0.3)
    A.set_coefficients(c1)                                   # update_coefficients! path
    _, B1 = make_pair("upwind", 1, 2, 0.1, n, c1)
    assert_close(G * u, O.apply_axis(B1, u, Qo), np.float64, "after update")
    # c*A (derivative_operator_functions.jl:165-197) and c*(L*Q) (ghost_derivative_operator.jl:66-76)
    _, B2 = make_pair("upwind", 1, 2, 0.1, n, c1)
    B2.scale(-2.5)
    assert_close((-2.5 * G) * u, O.apply_axis(B2, u, Qo), np.float64, "scalar * ghost")


def test_time_stepping_updates_coefficients_in_place(D, O):
    """A time-stepping loop with a time-dependent coefficient vector (update_coefficients!, abstract_operator_functions.jl:
    190-194): one cached plan, refreshed in place each step (no new plans, no device allocations), wind direction of the
    upwind rows re-selected from sign(c) every time, results equal to the oracle's at every step."""
    shape = (72, 40, 52)
    h = tuple(1.0 / (s + 1) for s in shape)
    dxs = [nonuniform_dx(s, hh, np.float64) for s, hh in zip(shape, h)]
    rows = np.arange(shape[2])
    c_of_t = lambda t: np.sin(0.37 * rows + t) * (1 + 0.5 * np.cos(3 * t))       # changes sign pattern from step to step
    Ux = D.UpwindDifference[3](1, 2, dxs[2], shape[2], c_of_t(0.0))
    A = D.CenteredDifference[1](2, 4, dxs[0], shape[0]) + D.CenteredDifference[2](2, 4, dxs[1], shape[1]) + Ux
    Q = D.compose(*D.RobinBC((1.0, 0.5, 0.25), (1.0, -0.5, 0.75), dxs, 1, shape))
    G = A * Q
    bcs = {ax + 1: O.RobinBC((1.0, 0.5, 0.25), (1.0, -0.5, 0.75), dxs[ax], 1) for ax in range(3)}
    u = uniform_field(shape, np.float64, seed=31)
    ud = D.DeviceArray.from_host(u)
    dud = D.DeviceArray(shape, np.float64)
    plans_seen = set()
    for step in range(6):
        t = 0.4 * step
        Ux.set_coefficients(c_of_t(t))
        D.mul_(dud, G, ud)
        cache = G.__dict__[D.apply._PLAN_CACHE_ATTR]
        assert len(cache) == 1
        plans_seen.add(id(next(iter(cache.values()))[0][0][0]))
        Bs = [O.CenteredDifference(2, 4, dxs[0], shape[0], axis=1), O.CenteredDifference(2, 4, dxs[1], shape[1], axis=2),
              O.UpwindDifference(1, 2, dxs[2], shape[2], c_of_t(t), axis=3)]
        assert_close(dud.to_host(), O.apply_sum(Bs, u, bcs), np.float64, f"step {step}")
    assert len(plans_seen) == 1, "coefficient updates must refresh the cached plan, not build new ones"


def test_host_buffer_path_equals_device_path(D):
    shape = (40, 36, 20)
    h = (0.1, 0.1, 0.1)
    A, _ = _laplacian_pair(shape, 4, h, np.float64)
    Q = D.compose(*D.RobinBC((1.0, 0.5, 0.25), (1.0, -0.5, 0.75), h, 1, shape))
    u = uniform_field(shape, np.float64, seed=101)
    host = (A * Q) * u                                        # numpy in -> numpy out (H2D + kernel + D2H)
    dev = ((A * Q) * D.DeviceArray.from_host(u)).to_host()
    assert isinstance(host, np.ndarray) and np.array_equal(host, dev)


def test_host_buffer_pipeline_equals_device_path(D, monkeypatch):
    """deo_plan_apply_host cuts the field into chunks of planes (upload | kernel | download overlapped); with a small
    chunk size every kernel family goes through several chunks and must reproduce the one-shot device result bitwise."""
    monkeypatch.setenv("DEO_HOST_CHUNK_BYTES", "4096")
    cases = []
    shape3, shape2 = (64, 48, 230), (128, 700)
    h3, h2 = tuple(1.0 / (s + 1) for s in shape3), tuple(1.0 / (s + 1) for s in shape2)
    lap3 = D.CenteredDifference[1](2, 4, h3[0], shape3[0]) + D.CenteredDifference[2](2, 4, h3[1], shape3[1]) + D.CenteredDifference[3](2, 4, h3[2], shape3[2])
    Q3 = D.compose(*D.RobinBC((1.0, 0.5, 0.25), (1.0, -0.5, 0.75), h3, 1, shape3))
    cases.append(("star", lap3 * Q3, shape3, 0))
    cases.append(("generic", lap3 * Q3, shape3, _flags(D, True)))
    c = np.sin(6 * np.pi * np.arange(1, shape3[2] + 1) / shape3[2])
    up = lap3 + D.UpwindDifference[3](1, 2, h3[2], shape3[2], c)
    cases.append(("star-table", up * Q3, shape3, 0))
    lap2 = D.CenteredDifference[1](2, 4, h2[0], shape2[0]) + D.CenteredDifference[2](2, 4, h2[1], shape2[1])
    Q2 = D.compose(*D.RobinBC((1.0, 0.5, 0.25), (1.0, -0.5, 0.75), h2, 1, shape2))
    cases.append(("star", lap2 * Q2, shape2, 0))
    for kern, G, shape, flags in cases:
        assert D.build_plans(G, shape, shape, np.float64, flags=flags)[0][0].info[0] == kern
        u = uniform_field(shape, np.float64, seed=len(shape))
        dev = D.DeviceArray(shape, np.float64)
        D.mul_(dev, G, D.DeviceArray.from_host(u), flags=flags)
        host = np.zeros(shape, order="F")
        D.mul_(host, G, u, flags=flags)
        assert np.array_equal(host, dev.to_host()), kern
        D.mul_(host, G, u, flags=flags)               # second call reuses the staging buffers
        assert np.array_equal(host, dev.to_host()), kern


def test_small_grids_and_errors(D, O):
    # BasicSDOExamples sizes: M = 3
    for kind, d, a in [("centered", 2, 2), ("upwind", 1, 1), ("upwind", 1, 2)]:
        for n in (3, 4, 7, 36, 37, 38):
            A, B = make_pair(kind, d, a, 0.25, n, -0.5 if kind == "upwind" else 1)
            Qd, Qo = bc_pair(("neumann", (0.3, -0.2), 1), 0.25, np.float64)
            u = uniform_field(n, np.float64, seed=n)
            assert_close((A * Qd) * u, O.apply_axis(B, u, Qo), np.float64, f"tiny {kind} n={n}")
    A, _ = make_pair("centered", 2, 4, 0.1, 20)
    with pytest.raises(AssertionError):
        D.mul_(np.zeros(20, order="F"), A, np.zeros(21))     # differentiated dimension must be padded by 2
    with pytest.raises(D.DeoError):
        A2, _ = make_pair("centered", 2, 4, 0.1, 21)
        D.mul_(np.zeros(20, order="F"), A2, np.zeros(22))    # len mismatch
    with pytest.raises(TypeError):
        A * np.zeros(22, dtype=np.int64)


def test_star_and_generic_kernels_agree_bitwise_where_both_apply(D):
    shape = (96, 40, 33)
    h = (0.1, 0.2, 0.3)
    A, _ = _laplacian_pair(shape, 6, h, np.float64)
    Q = D.compose(*D.RobinBC((1.0, 0.5, 0.25), (1.0, -0.5, 0.75), h, 1, shape))
    u = D.DeviceArray.from_host(uniform_field(shape, np.float64, seed=111))
    a = D.mul_alloc(A * Q, u).to_host()
    b = D.mul_alloc(A * Q, u, flags=D._lib.DEO_FLAG_FORCE_GENERIC).to_host()
    assert rel_err(a, b) <= 4e-16


# ------------------------------------------------------------------------------------------------------
# slab decomposition, single-process emulation (halo planes copied by hand)
# ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("nranks", [2, 3, 4])
def test_slab_plans_reassemble_the_global_result(D, O, nranks):
    import ctypes as C
    shape = (40, 24, 61)
    h = (0.1, 0.1, 0.1)
    for a in (2, 4, 6):
        A, Bs = _laplacian_pair(shape, a, h, np.float64)
        Q = D.compose(*D.RobinBC((1.0, 0.5, 0.25), (1.0, -0.5, 0.75), h, 1, shape))
        bcs = {ax: O.RobinBC((1.0, 0.5, 0.25), (1.0, -0.5, 0.75), h[ax - 1], 1) for ax in (1, 2, 3)}
        u = uniform_field(shape, np.float64, seed=121)
        want = O.apply_sum(Bs, u, bcs)
        terms = [(L.L, L.L.axis - 1) for L in (A * Q).ops]
        bclist = [D.apply._bc_for_axis(Q, ax, 3) for ax in (1, 2, 3)]
        out = np.zeros(shape, order="F")
        for r in range(nranks):
            plan = D.Plan(terms, bclist, shape, [False] * 3, np.float64, local_rank=(r, nranks))
            halo = C.c_int32(0)
            D._lib.check(D._lib.load().deo_dist_plan_halo(plan._h, C.byref(halo)))
            H = halo.value
            assert H == {2: 1, 4: 2, 6: 3}[a]
            s = C.c_int64(0); c = C.c_int64(0)
            D._lib.check(D._lib.load().deo_dist_slab(shape[2], nranks, r, C.byref(s), C.byref(c)))
            s, c = s.value, c.value
            ext = np.full((shape[0], shape[1], c + 2 * H), np.nan, order="F")     # NaN: a stray read of an unfilled halo shows
            lo, hi = max(s - H, 0), min(s + c + H, shape[2])
            ext[:, :, lo - (s - H):hi - (s - H)] = u[:, :, lo:hi]
            ud = D.DeviceArray.from_host(ext)
            dud = D.DeviceArray((shape[0], shape[1], c), np.float64)
            D._lib.check(D._lib.load().deo_dist_plan_apply(plan._h, dud._h, ud._h))
            out[:, :, s:s + c] = dud.to_host()
        assert_close(out, want, np.float64, f"slabs P={nranks} a={a}")
