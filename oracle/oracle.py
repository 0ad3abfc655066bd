"""TEST INFRASTRUCTURE -- NOT PRODUCT CODE.

ctypes front-end of the CPU oracle (oracle/deo_oracle.c), the C restatement of
the reference's operator-application path.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this module; the
product package (diffeqoperators.jl_b200) never does.

Arrays follow the reference's (Julia's) convention: shape (n1, n2, n3) with the
FIRST index fastest in memory (Fortran order); `axis` arguments are 1-based like
the `{N}` type parameter of CenteredDifference{N}.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libdeo_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("deo_oracle.c", "deo_oracle_impl.inc")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return so


class _Op64(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "derivative_order", "approximation_order", "len", "stencil_length",
        "boundary_stencil_length", "boundary_point_count", "offside", "wind",
        "nonuniform", "n_interior", "n_low", "n_high")] + [
        ("stencil_coefs", C.POINTER(C.c_double)), ("low_boundary_coefs", C.POINTER(C.c_double)),
        ("high_boundary_coefs", C.POINTER(C.c_double)), ("coefficients", C.POINTER(C.c_double))]


class _Op32(C.Structure):
    _fields_ = _Op64._fields_[:12] + [
        ("stencil_coefs", C.POINTER(C.c_float)), ("low_boundary_coefs", C.POINTER(C.c_float)),
        ("high_boundary_coefs", C.POINTER(C.c_float)), ("coefficients", C.POINTER(C.c_float))]


class _Bc64(C.Structure):
    _fields_ = [("kind", C.c_int), ("K_l", C.c_int), ("K_r", C.c_int), ("per_face", C.c_int),
                ("a_l", C.POINTER(C.c_double)), ("b_l", C.POINTER(C.c_double)),
                ("a_r", C.POINTER(C.c_double)), ("b_r", C.POINTER(C.c_double))]


class _Bc32(C.Structure):
    _fields_ = _Bc64._fields_[:4] + [
        ("a_l", C.POINTER(C.c_float)), ("b_l", C.POINTER(C.c_float)),
        ("a_r", C.POINTER(C.c_float)), ("b_r", C.POINTER(C.c_float))]


def _suffix(dtype):
    dtype = np.dtype(dtype)
    if dtype == np.float64:
        return "f64", C.c_double, _Op64, _Bc64
    if dtype == np.float32:
        return "f32", C.c_float, _Op32, _Bc32
    raise TypeError(f"oracle supports float32/float64, got {dtype}")


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        for suf, ct, Op, Bc in (_suffix(np.float64), _suffix(np.float32)):
            P = C.POINTER(ct)
            f = getattr(L, f"deo_oracle_calculate_weights_{suf}")
            f.argtypes = [C.c_int, ct, P, C.c_int, P]; f.restype = C.c_int
            f = getattr(L, f"deo_oracle_cumsum_{suf}")
            f.argtypes = [P, C.c_long, P]; f.restype = None
            for name, args in (
                ("centered_uniform", [C.c_int, C.c_int, ct, C.c_int, C.c_int, P]),
                ("centered_nonuniform", [C.c_int, C.c_int, P, C.c_int, C.c_int, P]),
                ("upwind_uniform", [C.c_int, C.c_int, ct, C.c_int, C.c_int, C.c_int, P]),
                ("upwind_nonuniform", [C.c_int, C.c_int, P, C.c_int, C.c_int, C.c_int, P]),
            ):
                f = getattr(L, f"deo_oracle_{name}_{suf}")
                f.argtypes = args; f.restype = C.POINTER(Op)
            f = getattr(L, f"deo_oracle_op_free_{suf}")
            f.argtypes = [C.POINTER(Op)]; f.restype = None
            f = getattr(L, f"deo_oracle_op_scale_{suf}")
            f.argtypes = [C.POINTER(Op), C.c_int, P]; f.restype = None
            f = getattr(L, f"deo_oracle_robin_uniform_{suf}")
            f.argtypes = [P, P, ct, C.c_int, P, P, P, P]; f.restype = C.c_int
            f = getattr(L, f"deo_oracle_robin_nonuniform_{suf}")
            f.argtypes = [P, P, P, C.c_int, C.c_int, P, P, P, P]; f.restype = C.c_int
            f = getattr(L, f"deo_oracle_general_{suf}")
            f.argtypes = [P, C.c_int, P, C.c_int, ct, P, C.c_int, C.c_int, P, P, P, P]; f.restype = C.c_int
            f = getattr(L, f"deo_oracle_mul1d_plain_{suf}")
            f.argtypes = [P, C.POINTER(Op), P, C.c_int]; f.restype = None
            f = getattr(L, f"deo_oracle_mul1d_bpv_{suf}")
            f.argtypes = [P, C.POINTER(Op), ct, ct, P, C.c_int]; f.restype = None
            f = getattr(L, f"deo_oracle_apply_axis_{suf}")
            f.argtypes = [C.POINTER(Op), C.c_int, C.c_int, C.POINTER(C.c_long), P, C.POINTER(Bc), P,
                          C.c_int, C.c_int]
            f.restype = C.c_int
        _LIB = L
    return _LIB


def _ptr(a, ct):
    return a.ctypes.data_as(C.POINTER(ct))


def calculate_weights(order, x0, x, dtype=np.float64):
    """fornberg.jl:7-62 in arithmetic type `dtype`."""
    suf, ct, _, _ = _suffix(dtype)
    x = np.ascontiguousarray(x, dtype=dtype)
    out = np.empty(len(x), dtype=dtype)
    rc = getattr(lib(), f"deo_oracle_calculate_weights_{suf}")(int(order), ct(x0), _ptr(x, ct), len(x), _ptr(out, ct))
    if rc:
        raise AssertionError("Not enough points for the requested order.")
    return out


def julia_cumsum(v, dtype=np.float64):
    suf, ct, _, _ = _suffix(dtype)
    v = np.ascontiguousarray(v, dtype=dtype)
    out = np.empty_like(v)
    getattr(lib(), f"deo_oracle_cumsum_{suf}")(_ptr(v, ct), len(v), _ptr(out, ct))
    return out


class Op:
    """Oracle-side DerivativeOperator (derivative_operator.jl:14-29)."""

    def __init__(self, ptr, dtype, axis):
        self._ptr = ptr
        self.dtype = np.dtype(dtype)
        self.axis = int(axis)            # 1-based, the {N} parameter
        self._suf, self._ct, _, _ = _suffix(dtype)

    def __del__(self):
        try:
            getattr(lib(), f"deo_oracle_op_free_{self._suf}")(self._ptr)
        except Exception:
            pass

    def __getattr__(self, name):
        s = self._ptr.contents
        if name in ("stencil_coefs", "low_boundary_coefs", "high_boundary_coefs", "coefficients"):
            sl, bsl = s.stencil_length, s.boundary_stencil_length
            sets = 2 if (s.wind and s.nonuniform) else 1
            shape = {
                "stencil_coefs": (sets, s.n_interior, sl),
                "low_boundary_coefs": (sets, s.n_low, bsl),
                "high_boundary_coefs": (sets, s.n_high, bsl),
                "coefficients": (s.len,),
            }[name]
            n = int(np.prod(shape))
            arr = np.ctypeslib.as_array(getattr(s, name), shape=(max(n, 1),))[:n].copy().reshape(shape)
            if name != "coefficients" and sets == 1:
                arr = arr[0]
                if name == "stencil_coefs" and not s.nonuniform:
                    arr = arr[0]
            return arr
        if name in dict(_Op64._fields_):
            return getattr(s, name)
        raise AttributeError(name)

    def scale(self, c):
        """c*A, derivative_operator_functions.jl:165-197 (in place on this copy)."""
        c = np.atleast_1d(np.asarray(c, dtype=self.dtype))
        getattr(lib(), f"deo_oracle_op_scale_{self._suf}")(self._ptr, int(c.size > 1), _ptr(np.ascontiguousarray(c), self._ct))
        return self

    # convolutions.jl:17-22 on a padded vector
    def mul1d_plain(self, x, y=None, overwrite=True):
        x = np.ascontiguousarray(x, dtype=self.dtype)
        assert x.size == self.len + 2
        y = np.zeros(self.len, dtype=self.dtype) if y is None else y
        getattr(lib(), f"deo_oracle_mul1d_plain_{self._suf}")(_ptr(y, self._ct), self._ptr, _ptr(x, self._ct), int(overwrite))
        return y

    def mul1d_bpv(self, l, r, u, y=None, overwrite=True):
        u = np.ascontiguousarray(u, dtype=self.dtype)
        assert u.size == self.len
        y = np.zeros(self.len, dtype=self.dtype) if y is None else y
        getattr(lib(), f"deo_oracle_mul1d_bpv_{self._suf}")(_ptr(y, self._ct), self._ptr, self._ct(l), self._ct(r), _ptr(u, self._ct), int(overwrite))
        return y

    def to_matrix(self):
        """convert_by_multiplication (test/DerivativeOperators/derivative_operators_interface.jl:102-117)."""
        n = self.len
        mat = np.zeros((n, n + 2), dtype=self.dtype)
        v = np.zeros(n + 2, dtype=self.dtype)
        for i in range(n + 2):
            v[i] = 1
            mat[:, i] = self.mul1d_plain(v)
            v[i] = 0
        return mat


def _coeff_arg(coeff, n, dtype):
    isvec = int(np.ndim(coeff) > 0)
    c = np.atleast_1d(np.asarray(coeff, dtype=dtype)).copy()
    if isvec and c.size != n:
        raise ValueError("coefficient vector length must equal len")
    return c, isvec


def CenteredDifference(d, a, dx, n, coeff=1, axis=1, dtype=np.float64):
    suf, ct, _, _ = _suffix(dtype)
    c, isvec = _coeff_arg(coeff, n, dtype)
    L = lib()
    if np.ndim(dx) == 0:
        p = getattr(L, f"deo_oracle_centered_uniform_{suf}")(d, a, ct(dx), n, isvec, _ptr(c, ct))
    else:
        dxv = np.ascontiguousarray(dx, dtype=dtype)
        assert dxv.size == n + 1, "dx must hold len+1 spacings"
        p = getattr(L, f"deo_oracle_centered_nonuniform_{suf}")(d, a, _ptr(dxv, ct), n, isvec, _ptr(c, ct))
    if not p:
        raise AssertionError("approximation_order must be greater than 1.")
    return Op(p, dtype, axis)


def UpwindDifference(d, a, dx, n, coeff=1, offside=0, axis=1, dtype=np.float64):
    suf, ct, _, _ = _suffix(dtype)
    c, isvec = _coeff_arg(coeff, n, dtype)
    L = lib()
    if np.ndim(dx) == 0:
        p = getattr(L, f"deo_oracle_upwind_uniform_{suf}")(d, a, ct(dx), n, offside, isvec, _ptr(c, ct))
    else:
        dxv = np.ascontiguousarray(dx, dtype=dtype)
        assert dxv.size == n + 1, "dx must hold len+1 spacings"
        p = getattr(L, f"deo_oracle_upwind_nonuniform_{suf}")(d, a, _ptr(dxv, ct), n, offside, isvec, _ptr(c, ct))
    if not p:
        raise AssertionError("invalid offside")
    return Op(p, dtype, axis)


class BC:
    """Affine atomic BC as the 4-tuple (a_l, b_l, a_r, b_r) (bc_operators.jl:21-25),
    or periodic.  For N-D use, `per_face` tables hold one BC per boundary pencil."""

    def __init__(self, a_l, b_l, a_r, b_r, dtype, kind=1):
        self.dtype = np.dtype(dtype)
        self.kind = kind
        self.a_l = np.ascontiguousarray(a_l, dtype=dtype)
        self.a_r = np.ascontiguousarray(a_r, dtype=dtype)
        self.b_l = np.ascontiguousarray(np.atleast_1d(b_l), dtype=dtype)
        self.b_r = np.ascontiguousarray(np.atleast_1d(b_r), dtype=dtype)
        self.per_face = int(self.a_l.ndim == 2)
        self.K_l = self.a_l.shape[-1] if kind == 1 else 0
        self.K_r = self.a_r.shape[-1] if kind == 1 else 0

    def cstruct(self):
        _, ct, _, Bc = _suffix(self.dtype)
        return Bc(self.kind, self.K_l, self.K_r, self.per_face, _ptr(self.a_l, ct), _ptr(self.b_l, ct),
                  _ptr(self.a_r, ct), _ptr(self.b_r, ct))

    def ghosts(self, u):
        """bc_operators.jl:188-192 on a vector -> (l, r)."""
        u = np.asarray(u, dtype=self.dtype)
        if self.kind == 2:
            return u[-1], u[0]
        dl = self.dtype.type(0)
        for k in range(self.K_l):
            dl = dl + self.a_l[k] * u[k]
        dr = self.dtype.type(0)
        for k in range(self.K_r):
            dr = dr + self.a_r[k] * u[len(u) - self.K_r + k]
        return dl + self.b_l[0], dr + self.b_r[0]


def RobinBC(l, r, dx, order=1, dtype=np.float64):
    suf, ct, _, _ = _suffix(dtype)
    order = int(order)
    la = np.asarray(l, dtype=dtype); ra = np.asarray(r, dtype=dtype)
    a_l = np.empty(order, dtype=dtype); a_r = np.empty(order, dtype=dtype)
    b_l = np.empty(1, dtype=dtype); b_r = np.empty(1, dtype=dtype)
    L = lib()
    with np.errstate(all="ignore"):
        if np.ndim(dx) == 0:
            rc = getattr(L, f"deo_oracle_robin_uniform_{suf}")(_ptr(la, ct), _ptr(ra, ct), ct(dx), order,
                                                               _ptr(a_l, ct), _ptr(b_l, ct), _ptr(a_r, ct), _ptr(b_r, ct))
        else:
            dxv = np.ascontiguousarray(dx, dtype=dtype)
            rc = getattr(L, f"deo_oracle_robin_nonuniform_{suf}")(_ptr(la, ct), _ptr(ra, ct), _ptr(dxv, ct), dxv.size, order,
                                                                  _ptr(a_l, ct), _ptr(b_l, ct), _ptr(a_r, ct), _ptr(b_r, ct))
    assert rc == 0
    return BC(a_l, b_l, a_r, b_r, dtype)


def NeumannBC(alpha, dx, order=1, dtype=np.float64):          # bc_operators.jl:173-175
    return RobinBC((0, 1, alpha[0]), (0, 1, alpha[1]), dx, order, dtype)


def DirichletBC(al, ar, dtype=np.float64):                     # bc_operators.jl:176-178
    return RobinBC((1, 0, al), (1, 0, ar), 1, 2, dtype)


def Dirichlet0BC(dtype=np.float64):                            # bc_operators.jl:180
    return DirichletBC(0, 0, dtype)


def Neumann0BC(dx, order=1, dtype=np.float64):                 # bc_operators.jl:181-183
    return NeumannBC((0, 0), dx, order, dtype)


def PeriodicBC(dtype=np.float64):                              # bc_operators.jl:17-19
    z = np.zeros(0, dtype=dtype)
    return BC(z, 0, z, 0, dtype, kind=2)


def GeneralBC(alpha_l, alpha_r, dx, order=1, dtype=np.float64):
    suf, ct, _, _ = _suffix(dtype)
    order = int(order)
    al = np.ascontiguousarray(alpha_l, dtype=dtype); ar = np.ascontiguousarray(alpha_r, dtype=dtype)
    a_l = np.empty(order + al.size - 3, dtype=dtype); a_r = np.empty(order + ar.size - 3, dtype=dtype)
    b_l = np.empty(1, dtype=dtype); b_r = np.empty(1, dtype=dtype)
    if np.ndim(dx) == 0:
        dxs, dxv, ndx = ct(dx), None, 0
    else:
        dxa = np.ascontiguousarray(dx, dtype=dtype)
        dxs, dxv, ndx = ct(0), _ptr(dxa, ct), dxa.size
    rc = getattr(lib(), f"deo_oracle_general_{suf}")(_ptr(al, ct), al.size, _ptr(ar, ct), ar.size, dxs, dxv, ndx, order,
                                                      _ptr(a_l, ct), _ptr(b_l, ct), _ptr(a_r, ct), _ptr(b_r, ct))
    assert rc == 0
    return BC(a_l, b_l, a_r, b_r, dtype)


def _collapse(shape, axis0):
    """N-D -> (pre, n, post): derivative_operator_functions.jl:27-69 only ever
    walks pencils along one axis, so any N collapses to three dims."""
    pre = int(np.prod(shape[:axis0], dtype=np.int64))
    post = int(np.prod(shape[axis0 + 1:], dtype=np.int64))
    return pre, int(shape[axis0]), post


def apply_axis(op: Op, u, bc: BC | None = None, out=None, overwrite=True, nthreads=1):
    """One operator along op.axis.
    bc is None : `u` is a pre-padded array M; dims other than op.axis may carry
                 unnecessary padding, which is stripped (2:end-1) exactly as
                 derivative_operator_functions.jl:50-57 does when `out` fixes the
                 output shape (default: only the op axis is treated as padded).
    bc given   : `u` is the unpadded field, ghosts come from Q*u."""
    dt = op.dtype
    u = np.asarray(u, dtype=dt)
    ax = op.axis - 1
    if bc is None:
        oshape = list(u.shape); oshape[ax] -= 2
        if out is not None:
            oshape = list(out.shape)
        sl = tuple(slice(None) if (d == ax or u.shape[d] == oshape[d]) else slice(1, -1) for d in range(u.ndim))
        u = u[sl]
    else:
        oshape = list(u.shape)
    uf = np.asfortranarray(u)
    res = np.zeros(oshape, dtype=dt, order="F") if out is None else np.asfortranarray(out).copy(order="F")
    pre, n, post = _collapse(oshape, ax)
    ndims_eff = 1 if u.ndim == 1 else 3
    dims = (C.c_long * 3)(*((n, 1, 1) if ndims_eff == 1 else (pre, n, post)))
    axis_eff = 0 if ndims_eff == 1 else 1
    _, ct, _, Bc = _suffix(dt)
    bcs = bc.cstruct() if bc is not None else None
    rc = getattr(lib(), f"deo_oracle_apply_axis_{op._suf}")(
        op._ptr, axis_eff, ndims_eff, dims, uf.ctypes.data_as(C.POINTER(ct)),
        C.byref(bcs) if bcs is not None else None, res.ctypes.data_as(C.POINTER(ct)), int(overwrite), int(nthreads))
    if rc:
        raise RuntimeError(f"oracle apply_axis failed rc={rc}")
    if out is not None:
        out[...] = res
        return out
    return res


def apply_sum(ops, u, bcs=None, nthreads=1):
    """composite_operators.jl:64-65: sum(op -> op*x, ops), left fold.
    bcs: dict {axis(1-based): BC} (the composed Q) or None for a pre-padded
    array (padded on every axis that carries an op; SURVEY 2.1-2 intended
    semantics: perpendicular padding stripped with 2:end-1)."""
    u = np.asarray(u)
    acc = None
    if bcs is None:
        padded_axes = sorted({op.axis - 1 for op in ops})
        oshape = tuple(s - 2 if d in padded_axes else s for d, s in enumerate(u.shape))
    for op in ops:
        if bcs is None:
            r = apply_axis(op, u, None, out=np.zeros(oshape, dtype=op.dtype, order="F"), nthreads=nthreads)
        else:
            r = apply_axis(op, u, bcs[op.axis], nthreads=nthreads)
        acc = r if acc is None else acc + r
    return acc
