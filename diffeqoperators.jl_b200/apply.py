"""`mul!` / `*` of the mirrored operator objects: walks the operator (A.ops, A.L, A.Q), hands the
fields to libdeo_b200 as a deo_plan_desc (cached per operator/shape), and applies it.

Replaces, call for call:
  convolutions.jl:17-22                          mul!(y, A::DerivativeOperator, x::AbstractVector)
  derivative_operator_functions.jl:18-69,150-163 mul!/ * on N-D arrays (pre-padded input)
  ghost_derivative_operator.jl:15-37             mul!/ * of L*Q
  composite_operators.jl:64-65,76-83             mul!/ * of sums
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .bc import (AffineBC, AtomicBC, BoundaryPadded, ComposedMultiDimBC, MultiDimDirectionalBC, PeriodicBC)
from .device import DeviceArray
from .operators import DerivativeOperator, DiffEqOperatorCombination, GhostDerivativeOperator


# ---- operator normalisation ------------------------------------------------------------------------
def _terms(A):
    """-> [(DerivativeOperator, Q or None)] in application order."""
    if isinstance(A, DerivativeOperator):
        return [(A, None)]
    if isinstance(A, GhostDerivativeOperator):
        L = A.L
        if isinstance(L, DiffEqOperatorCombination):
            return [(op, A.Q) for op, _ in sum((_terms(o) for o in L.ops), [])]
        return [(L, A.Q)]
    if isinstance(A, DiffEqOperatorCombination):
        return sum((_terms(op) for op in A.ops), [])
    raise TypeError(f"cannot apply {type(A).__name__}")


def _bc_for_axis(Q, axis, ndims):
    """The boundary operator of `Q` that extends dimension `axis` (1-based) of an ndims-array."""
    if Q is None:
        return None
    if isinstance(Q, AtomicBC):
        # vector: Q itself; array: MultiDimBC{1}(Q, size(u))  (multi_dim_bc_operators.jl:210)
        return Q if axis == 1 else None
    if isinstance(Q, MultiDimDirectionalBC):
        return Q.BCs if Q.axis == axis else None
    if isinstance(Q, ComposedMultiDimBC):
        return Q.parts[axis - 1].BCs
    raise TypeError(f"unknown boundary operator {type(Q).__name__}")


class _Keep:
    """Keeps numpy arrays referenced by a descriptor alive until deo_plan_create has copied them."""

    def __init__(self):
        self.refs = []

    def ptr(self, a, dtype):
        a = np.ascontiguousarray(a, dtype=dtype)
        self.refs.append(a)
        return a.ctypes.data_as(C.c_void_p)


def _op_desc(op: DerivativeOperator, axis0: int, keep: _Keep) -> _lib.OpDesc:
    T = op.T
    d = _lib.OpDesc()
    d.axis = axis0
    d.kind = _lib.DEO_OP_UPWIND if op.wind else _lib.DEO_OP_CENTERED
    d.nonuniform = int(op.nonuniform)
    d.derivative_order = op.derivative_order
    d.len = op.len
    d.stencil_length = op.stencil_length
    d.boundary_stencil_length = op.boundary_stencil_length
    d.boundary_point_count = op.boundary_point_count
    d.offside = op.offside
    d.stencil_coefs = keep.ptr(op.stencil_coefs, T)
    d.low_boundary_coefs = keep.ptr(op.low_boundary_coefs, T) if op.low_boundary_coefs.size else None
    d.high_boundary_coefs = keep.ptr(op.high_boundary_coefs, T) if op.high_boundary_coefs.size else None
    d.coefficients = keep.ptr(op.coefficients, T)
    return d


def _bc_desc(bc, T, keep: _Keep, ndims: int = 1) -> _lib.BcDesc:
    d = _lib.BcDesc()
    if bc is None:
        d.kind = _lib.DEO_BC_NONE
        return d
    if isinstance(bc, PeriodicBC):
        if ndims > 1:
            # On arrays the reference's periodic ghosts are lower = u[1, ...], upper = u[end, ...] of the SAME pencil
            # (multi_dim_bc_operators.jl:221-228) -- literally the affine BC a = [1], b = 0, which every
            # kernel (tiled ones included) evaluates exactly: fma(1, u, 0) + 0 == u.
            one, zero = np.ones(1, dtype=T), np.zeros(1, dtype=T)
            d.kind = _lib.DEO_BC_AFFINE
            d.per_face = 0
            d.K_l = d.K_r = 1
            d.a_l, d.a_r = keep.ptr(one, T), keep.ptr(one, T)
            d.b_l, d.b_r = keep.ptr(zero, T), keep.ptr(zero, T)
            return d
        d.kind = _lib.DEO_BC_PERIODIC          # vectors: l = u[end], r = u[1] (bc_operators.jl:192): the wrap-around read
        return d
    if isinstance(bc, AffineBC):
        d.kind = _lib.DEO_BC_AFFINE
        d.per_face = 0
        d.K_l, d.K_r = bc.a_l.size, bc.a_r.size
        d.a_l = keep.ptr(bc.a_l, T) if bc.a_l.size else None
        d.a_r = keep.ptr(bc.a_r, T) if bc.a_r.size else None
        d.b_l = keep.ptr(np.array([bc.b_l]), T)
        d.b_r = keep.ptr(np.array([bc.b_r]), T)
        return d
    if isinstance(bc, np.ndarray):       # one atomic BC per boundary pencil (MultiDimDirectionalBC.BCs)
        flat = bc.reshape(-1, order="F")
        if not all(isinstance(q, AffineBC) for q in flat):
            raise NotImplementedError("per-pencil BC arrays must hold affine BCs (Robin/General/Dirichlet/Neumann)")
        Kl = max(q.a_l.size for q in flat)
        Kr = max(q.a_r.size for q in flat)
        a_l = np.zeros((flat.size, Kl), dtype=T)
        a_r = np.zeros((flat.size, Kr), dtype=T)
        for f, q in enumerate(flat):
            a_l[f, :q.a_l.size] = q.a_l                  # a_l . u[1:K]: shorter stencils padded at the far end
            a_r[f, Kr - q.a_r.size:] = q.a_r             # a_r . u[end-K+1:end]: padded at the near end
        d.kind = _lib.DEO_BC_AFFINE
        d.per_face = 1
        d.K_l, d.K_r = Kl, Kr
        d.a_l = keep.ptr(a_l, T) if Kl else None
        d.a_r = keep.ptr(a_r, T) if Kr else None
        d.b_l = keep.ptr(np.array([q.b_l for q in flat]), T)
        d.b_r = keep.ptr(np.array([q.b_r for q in flat]), T)
        return d
    raise TypeError(f"unsupported boundary condition {type(bc).__name__}")


class Plan:
    """Owns a deo_plan."""

    def __init__(self, ops_axes, bcs, out_shape, padded, dtype, accumulate=False, flags=0,
                 dist=None, local_rank=None):
        keep = _Keep()
        nd = len(out_shape)
        desc = _lib.PlanDesc()
        desc.dtype = _lib.dtype_code(dtype)
        desc.ndims = nd
        for a in range(_lib.DEO_MAX_DIMS):
            desc.dims[a] = out_shape[a] if a < nd else 1
            desc.padded[a] = int(padded[a]) if a < nd else 0
        arr = (_lib.OpDesc * len(ops_axes))(*[_op_desc(op, ax, keep) for op, ax in ops_axes])
        desc.nops = len(ops_axes)
        desc.ops = arr
        desc.accumulate = int(accumulate)
        desc.flags = flags
        for a in range(nd):
            desc.bc[a] = _bc_desc(bcs[a], dtype, keep, nd)
        h = C.c_void_p()
        L = _lib.load()
        if dist is not None:
            _lib.check(L.deo_dist_plan_create(dist, C.byref(desc), C.byref(h)))
        elif local_rank is not None:
            _lib.check(L.deo_dist_plan_create_local(C.byref(desc), local_rank[0], local_rank[1], C.byref(h)))
        else:
            _lib.check(L.deo_plan_create(C.byref(desc), C.byref(h)))
        self._h = h
        self.out_shape = tuple(out_shape)
        self.in_shape = tuple(s + 2 * int(p) for s, p in zip(out_shape, padded))
        self.dtype = np.dtype(dtype)
        self.term_index = list(range(len(ops_axes)))      # plan op k  <->  term of the operator it was built from

    def __del__(self):
        try:
            if self._h is not None:
                _lib.load().deo_plan_destroy(self._h)
                self._h = None
        except Exception:
            pass

    @property
    def info(self):
        name = C.create_string_buffer(64)
        n = C.c_int32(0)
        _lib.check(_lib.load().deo_plan_info(self._h, name, 64, C.byref(n)))
        return name.value.decode(), int(n.value)

    def apply(self, du: DeviceArray, u: DeviceArray):
        _lib.check(_lib.load().deo_plan_apply(self._h, du._h, u._h))

    def apply_axpy(self, out: DeviceArray, u: DeviceArray, dt: float):
        """out = u + dt * (A u), fused into the kernel's store where the plan runs on a tiled kernel."""
        _lib.check(_lib.load().deo_plan_apply_axpy(self._h, out._h, u._h, float(dt)))

    def apply_n(self, du: DeviceArray, u: DeviceArray, reps: int):
        _lib.check(_lib.load().deo_plan_apply_n(self._h, du._h, u._h, reps))

    def apply_host(self, du: np.ndarray, u: np.ndarray):
        _lib.check(_lib.load().deo_plan_apply_host(self._h, du.ctypes.data_as(C.c_void_p), u.ctypes.data_as(C.c_void_p)))

    def time(self, du: DeviceArray, u: DeviceArray, reps: int) -> float:
        ms = C.c_float(0)
        _lib.check(_lib.load().deo_plan_time(self._h, du._h, u._h, reps, C.byref(ms)))
        return float(ms.value)

    def update_coefficients(self, op_index: int, coeffs):
        c = np.ascontiguousarray(coeffs, dtype=self.dtype)
        _lib.check(_lib.load().deo_plan_update_coefficients(self._h, op_index, c.ctypes.data_as(C.c_void_p)))


def _collapse(shape, axis0):
    pre = int(np.prod(shape[:axis0], dtype=np.int64))
    post = int(np.prod(shape[axis0 + 1:], dtype=np.int64))
    return (pre, int(shape[axis0]), post)


def build_plans(A, out_shape, in_shape, dtype, accumulate=False, flags=0, Q_override=None):
    """-> list of (Plan, reshape_out, reshape_in); more than one only for >3-D arrays with operators on
    several axes (each axis then runs as its own (pre, n, post) pass, accumulating)."""
    terms = _terms(A)
    if Q_override is not None:
        terms = [(L, Q_override if Q is None else Q) for L, Q in terms]
    nd = len(out_shape)
    T = np.dtype(dtype)
    for L, _ in terms:
        if L.T != T:
            raise TypeError(f"eltype(A) = {L.T} must equal eltype(u) = {T} on the device path")
        if L.axis > nd:
            raise AssertionError(f"operator axis {L.axis} exceeds ndims(u) = {nd}")
    padded = [int(i) - int(o) for i, o in zip(in_shape, out_shape)]
    if any(p not in (0, 2) for p in padded):
        raise AssertionError(f"size(M) = {tuple(in_shape)} must equal size(x_temp) = {tuple(out_shape)} or exceed it by 2 per dimension")
    padded = [p == 2 for p in padded]

    def bcs_for(term_list, axes_1based):
        out = {}
        for L, Q in term_list:
            ax = L.axis
            bc = None if padded[ax - 1] else _bc_for_axis(Q, ax, nd)
            if not padded[ax - 1] and bc is None:
                raise AssertionError("the differentiated dimension must be padded: size(x_temp, N) + 2 == size(M, N), or carry a boundary condition")
            if ax in out and out[ax] is not bc:
                same = isinstance(bc, AffineBC) and isinstance(out[ax], AffineBC) and \
                    np.array_equal(bc.a_l, out[ax].a_l) and np.array_equal(bc.a_r, out[ax].a_r) and bc.b_l == out[ax].b_l and bc.b_r == out[ax].b_r
                if not same:
                    raise NotImplementedError("operators on one axis with different boundary conditions must be applied separately")
            out[ax] = bc
        return [out.get(a) for a in axes_1based]

    if nd <= 3:
        bcs = bcs_for(terms, range(1, nd + 1))
        plan = Plan([(L, L.axis - 1) for L, _ in terms], bcs, out_shape, padded, T, accumulate, flags)
        return [(plan, tuple(out_shape), tuple(in_shape))]
    # >3-D: one collapsed pass per axis (A.5b: an axis-k op on an N-D array is (prod dims[:k], n, prod dims[k+1:]))
    if any(padded[d] for d in range(nd)) and len({L.axis for L, _ in terms}) > 1:
        raise NotImplementedError(">3-D pre-padded input with operators on several axes")
    plans = []
    for ax in sorted({L.axis for L, _ in terms}, key=lambda a: [L.axis for L, _ in terms].index(a)):
        sub = [(L, Q) for L, Q in terms if L.axis == ax]
        bc = bcs_for(sub, [ax])[0]
        if isinstance(bc, np.ndarray):
            raise NotImplementedError("per-pencil BC arrays on >3-D input")
        # other padded dims are stripped on the host side by the caller (views), so only `ax` may be padded here
        cshape = _collapse(out_shape, ax - 1)
        cpad = [False, padded[ax - 1], False]
        cin = tuple(s + 2 * int(p) for s, p in zip(cshape, cpad))
        plan = Plan([(L, 1) for L, _ in sub], [None, bc, None], cshape, cpad, T, accumulate or bool(plans), flags)
        plan.term_index = [k for k, (L, _) in enumerate(terms) if L.axis == ax]
        plans.append((plan, cshape, cin))
    return plans


_PLAN_CACHE_ATTR = "_deo_plan_cache"
_PLAN_CACHE_MAX = 8          # entries per operator object (shape / dtype / boundary-operator variants); least recently used goes first


def _coeff_versions(A):
    return tuple(L._version for L, _ in _terms(A))


def _get_plans(A, out_shape, in_shape, dtype, accumulate, flags, Q_override=None):
    """Cached plans of operator `A` for one (shapes, dtype, accumulate, boundary operator).  Coefficient changes
    (`update_coefficients!`, `set_coefficients`, a time-dependent `coeff_func` in `A(du,u,p,t)`) do NOT create new plans:
    the cached plan is refreshed in place with deo_plan_update_coefficients for the operators whose version moved, so a
    time-stepping loop neither leaks device tables nor pays a plan build per step."""
    cache = A.__dict__.setdefault(_PLAN_CACHE_ATTR, {})
    key = (tuple(out_shape), tuple(in_shape), np.dtype(dtype).str, bool(accumulate), flags, id(Q_override))
    entry = cache.pop(key, None)
    versions = _coeff_versions(A)
    if entry is None:
        # the entry keeps Q_override alive, so its id() cannot be recycled for another boundary operator while cached
        entry = [build_plans(A, out_shape, in_shape, dtype, accumulate, flags, Q_override), Q_override, versions]
    elif entry[2] != versions:
        terms = _terms(A)
        for plan, _, _ in entry[0]:
            for k, op_index in enumerate(plan.term_index):
                if entry[2][op_index] != versions[op_index]:
                    plan.update_coefficients(k, terms[op_index][0].coefficients)
        entry[2] = versions
    cache[key] = entry                      # (re)inserted last: dict order is the LRU order
    while len(cache) > _PLAN_CACHE_MAX:
        cache.pop(next(iter(cache)))
    return entry[0]


def _unwrap(u):
    """-> (array, Q_override).  `L * (Q*u)` is the same fused application as `(L*Q) * u`."""
    if isinstance(u, BoundaryPadded):
        return u.u, u.Q
    return u, None


def _out_shape_alloc(A, u_shape, has_bc):
    if has_bc:
        return tuple(u_shape)                      # unpadded_size(u)  (ghost_derivative_operator.jl:28,:34)
    axes = {L.axis for L, _ in _terms(A)}
    if len(axes) != 1:
        raise AssertionError("`*` of a sum of operators along different axes needs boundary conditions (use mul_ with an explicit du for pre-padded input)")
    s = list(u_shape)
    s[axes.pop() - 1] -= 2                         # _mul: size_x_temp[N] -= 2  (derivative_operator_functions.jl:157-163)
    return tuple(s)


def mul_(du, A, u, *, overwrite=True, flags=0):
    """LinearAlgebra.mul!(du, A, u).  du/u: both DeviceArray (no host traffic) or both numpy arrays
    (host-buffer path: one H2D copy of u, the fused kernel, one D2H copy of du)."""
    u, Qo = _unwrap(u)
    terms = _terms(A)
    if Qo is not None and any(Q is not None for _, Q in terms):
        raise TypeError("operator already carries a boundary condition")
    plans = _get_plans(A, du.shape, u.shape, du.dtype, not overwrite, flags, Qo)
    if isinstance(du, DeviceArray) != isinstance(u, DeviceArray):
        raise TypeError("du and u must both be DeviceArray or both be host arrays")
    if np.dtype(u.dtype) != np.dtype(du.dtype):
        raise TypeError("eltype(du) must equal eltype(u)")
    if isinstance(du, DeviceArray):
        for plan, _, _ in plans:
            plan.apply(du, u)
        return du
    if not (isinstance(du, np.ndarray) and du.flags.f_contiguous and du.flags.writeable):
        raise TypeError("host du must be a writeable column-major (Fortran-ordered) numpy array")
    uf = np.asfortranarray(u, dtype=du.dtype)
    for i, (plan, _, _) in enumerate(plans):
        plan.apply_host(du, uf)
    return du


def step_(out, A, u, dt, *, flags=0):
    """out = u + dt * (A*u): one explicit-Euler update with the AXPY fused into the operator application (the caller of
    mul! in the reference's own examples, test/DerivativeOperators/3D_laplacian.jl:20-24).  Device arrays only."""
    if not (isinstance(out, DeviceArray) and isinstance(u, DeviceArray)):
        raise TypeError("step_ works on DeviceArray operands")
    if out.shape != u.shape or out.dtype != u.dtype:
        raise TypeError("out must have the shape and eltype of u")
    plans = _get_plans(A, out.shape, u.shape, out.dtype, False, flags, None)
    if len(plans) != 1:
        raise NotImplementedError("step_ on > 3-D arrays with operators along several axes")
    plans[0][0].apply_axpy(out, u, dt)
    return out


def mul_alloc(A, u, *, flags=0):
    """Base.:*(A, u): allocates the output like the reference (`zeros(promote_type(...), unpadded_size(u))`)."""
    arr, Qo = _unwrap(u)
    has_bc = Qo is not None or any(Q is not None for _, Q in _terms(A))
    out_shape = _out_shape_alloc(A, arr.shape, has_bc)
    if isinstance(arr, DeviceArray):
        du = DeviceArray(out_shape, arr.dtype)
    else:
        arr = np.asarray(arr)
        if arr.dtype not in (np.dtype(np.float32), np.dtype(np.float64)):
            raise TypeError("the device path requires Float32 or Float64 arrays")
        du = np.zeros(out_shape, dtype=arr.dtype, order="F")
    return mul_(du, A, BoundaryPadded(Qo, arr) if Qo is not None else arr, flags=flags)


# ---- host-side Array(Q*u) (inspection / tests only; the kernels never materialise this) -----------------
def _ghosts_1d(bc, v):
    if isinstance(bc, PeriodicBC):
        return v[-1], v[0]
    T = bc.T.type
    dl = T(0)
    for k in range(bc.a_l.size):
        dl = dl + bc.a_l[k] * v[k]
    dr = T(0)
    for k in range(bc.a_r.size):
        dr = dr + bc.a_r[k] * v[len(v) - bc.a_r.size + k]
    return dl + bc.b_l, dr + bc.b_r


def materialize_padded(P: BoundaryPadded):
    u = np.asarray(P.u.to_host() if isinstance(P.u, DeviceArray) else P.u)
    Q = P.Q
    nd = u.ndim
    if nd == 1:
        l, r = _ghosts_1d(Q, u)
        return np.concatenate([[l], u, [r]]).astype(u.dtype)
    axes = P.padded_axes
    shape = [s + (2 if d + 1 in axes else 0) for d, s in enumerate(u.shape)]
    out = np.zeros(shape, dtype=u.dtype, order="F")            # corners/edges read as 0 (boundary_padded_arrays.jl:165-166)
    inner = tuple(slice(1, -1) if d + 1 in axes else slice(None) for d in range(nd))
    out[inner] = u
    for ax in axes:
        bcs = _bc_for_axis(Q, ax, nd)
        um = np.moveaxis(u, ax - 1, 0)
        om = np.moveaxis(out, ax - 1, 0)
        other = tuple(slice(1, -1) if (d + 1 in axes) else slice(None) for d in range(nd) if d != ax - 1)
        lo = np.zeros(um.shape[1:], dtype=u.dtype)
        hi = np.zeros(um.shape[1:], dtype=u.dtype)
        # The reference swaps the periodic ghosts (lower = u[1], upper = u[end]) only in the methods specialised on an array
        # whose element type is PeriodicBC itself (multi_dim_bc_operators.jl:221-228, :244-252).  A per-pencil array that
        # mixes PeriodicBC with other atomic BCs goes through the generic slice_rmul (:27-52), pencil by pencil, with the
        # 1-D rule lower = u[end], upper = u[1] (bc_operators.jl:192; test/DerivativeOperators/multi_dim_bc_test.jl:29-31).
        all_periodic = isinstance(bcs, PeriodicBC) or (isinstance(bcs, np.ndarray) and all(isinstance(q, PeriodicBC) for q in bcs.reshape(-1)))
        for idx in np.ndindex(*um.shape[1:]):
            bc = bcs if isinstance(bcs, AtomicBC) else bcs[idx]
            if all_periodic:
                lo[idx], hi[idx] = um[(0,) + idx], um[(-1,) + idx]
            else:
                lo[idx], hi[idx] = _ghosts_1d(bc, um[(slice(None),) + idx])
        om[(0,) + other] = lo
        om[(-1,) + other] = hi
    return out
