"""Host-side mirror of the reference's operator constructors (they stay on the host in the Julia
deployment too; only their *fields* cross the C ABI):

  fornberg.jl:7-62                 calculate_weights
  derivative_operator.jl:14-29     DerivativeOperator (same field names)
  derivative_operator.jl:81-205    CenteredDifference{N}  (uniform / non-uniform)
  derivative_operator.jl:454-639   UpwindDifference{N}    (uniform / non-uniform)
  coefficient_functions.jl:7-26    compute_coeffs!
  derivative_operator_functions.jl:165-197   c*A
  composite_operators.jl:15-83     DiffEqOperatorCombination (+)
  ghost_derivative_operator.jl     GhostDerivativeOperator (L*Q)

Julia's `CenteredDifference{N}(...)` is spelled `CenteredDifference[N](...)`; axes are 1-based.
Arithmetic is carried out in the element type T exactly where the reference does (numpy scalars).
"""
from __future__ import annotations

import numbers

import numpy as np


def _T(dtype):
    dt = np.dtype(dtype)
    if dt not in (np.dtype(np.float32), np.dtype(np.float64)):
        raise TypeError(f"element type must be Float32 or Float64, got {dt}")
    return dt.type


def _infer_dtype(dx, dtype):
    if dtype is not None:
        return np.dtype(dtype)
    if isinstance(dx, np.ndarray) or isinstance(dx, np.generic):
        dt = np.asarray(dx).dtype
        return dt if dt in (np.dtype(np.float32),) else np.dtype(np.float64)
    return np.dtype(np.float64)


def calculate_weights(order: int, x0, x, dtype=np.float64):
    """fornberg.jl:7-62.  Arithmetic type = typeof(x0) = `dtype`; includes the sum-to-zero fix
    `_C[div(N,2)+1] -= sum(_C)` (:57-61)."""
    T = _T(dtype)
    x = [T(v) for v in x]
    x0 = T(x0)
    N = len(x)
    assert order < N, "Not enough points for the requested order."
    M = order
    c1 = T(1)
    c4 = x[0] - x0
    Cm = [[T(0)] * (M + 1) for _ in range(N)]
    Cm[0][0] = T(1)
    with np.errstate(all="ignore"):
        for i in range(1, N):
            mn = min(i, M)
            c2 = T(1)
            c5 = c4
            c4 = x[i] - x0
            for j in range(i):
                c3 = x[i] - x[j]
                c2 = c2 * c3
                if j == i - 1:
                    for s in range(mn, 0, -1):
                        Cm[i][s] = c1 * (T(s) * Cm[i - 1][s - 1] - c5 * Cm[i - 1][s]) / c2
                    Cm[i][0] = -c1 * c5 * Cm[i - 1][0] / c2
                for s in range(mn, 0, -1):
                    Cm[j][s] = (c4 * Cm[j][s] - T(s) * Cm[j][s - 1]) / c3
                Cm[j][0] = c4 * Cm[j][0] / c3
            c1 = c2
        out = [Cm[i][M] for i in range(N)]
        if order != 0:
            s = out[0]
            for v in out[1:]:
                s = s + v
            out[N // 2] = out[N // 2] - s
    return np.array(out, dtype=dtype)


def julia_cumsum(v, dtype):
    """Base.cumsum on a Vector = accumulate_pairwise! (block size 128)."""
    T = _T(dtype)
    v = [T(a) for a in v]
    n = len(v)
    c = [T(0)] * n
    if n == 0:
        return np.array(c, dtype=dtype)

    def rec(s, i1, m):
        if m < 128:
            s_ = v[i1]
            c[i1] = s + s_
            for i in range(i1 + 1, i1 + m):
                s_ = s_ + v[i]
                c[i] = s + s_
            return s_
        n2 = m >> 1
        s_ = rec(s, i1, n2)
        s_ = s_ + rec(s + s_, i1 + n2, m - n2)
        return s_

    c[0] = v[0]
    if n > 1:
        rec(c[0], 1, n - 1)
    return np.array(c, dtype=dtype)


def compute_coeffs(coeff_func, n, dtype):
    """compute_coeffs!(coeff_func, zeros(T, len)), coefficient_functions.jl:7-26."""
    cur = np.zeros(n, dtype=dtype)
    if callable(coeff_func):
        try:
            res = np.asarray(coeff_func(cur), dtype=dtype)       # method for Vector{T} (:20-21)
            if res.shape != (n,):
                raise TypeError
            return res.copy()
        except TypeError:
            return np.array([coeff_func(v) for v in cur], dtype=dtype)   # map! (:23)
    if np.ndim(coeff_func) == 0:
        return cur + np.dtype(dtype).type(coeff_func)
    c = np.asarray(coeff_func, dtype=dtype)
    if c.shape != (n,):
        raise ValueError(f"coefficient vector has length {c.size}, operator len is {n}")
    return cur + c


class _Indexable(type):
    """`Ctor[N](...)` == Julia's `Ctor{N}(...)`."""

    def __getitem__(cls, axis):
        axis = int(axis)
        if axis < 1:
            raise ValueError("axis parameter is 1-based")
        return lambda *a, **k: cls(*a, axis=axis, **k)


class AbstractDiffEqLinearOperator:
    __array_ufunc__ = None      # numpy arrays on the left defer to __rmul__ (c .* A, not elementwise broadcasting)

    def __add__(self, other):
        return DiffEqOperatorCombination(_ops_of(self) + _ops_of(other))

    def __sub__(self, other):
        return self + (-other)

    def __call__(self, *args):
        # (L)(u,p,t) = L*u ; (L)(du,u,p,t) = mul!(du,L,u)   (src/DiffEqOperators.jl:53-54, :66-75)
        from .apply import mul_
        if len(args) == 3:
            u, p, t = args
            self.update_coefficients_(u, p, t)
            return self * u
        du, u, p, t = args
        self.update_coefficients_(u, p, t)
        return mul_(du, self, u)

    def update_coefficients_(self, u, p, t):
        pass


def _compose(L1, L2):
    """L1 * L2 of two operators = DiffEqOperatorComposition((L2, L1))  (composite_operators.jl:107-115)."""
    from .vector_calculus import compose_operators
    return compose_operators(L1, L2)


def _base(op):
    return op.L if isinstance(op, GhostDerivativeOperator) else op


def _ops_of(x):
    if isinstance(x, DiffEqOperatorCombination):
        return tuple(x.ops)
    if isinstance(x, AbstractDiffEqLinearOperator):
        return (x,)
    raise TypeError(f"cannot combine an operator with {type(x).__name__}")


class DerivativeOperator(AbstractDiffEqLinearOperator):
    """derivative_operator.jl:14-29 -- a dumb data structure holding the stencils.

    stencil_coefs        uniform: (sl,) ; non-uniform centered: (len-2bpc, sl) ;
                         non-uniform upwind: (2, len-2bpc, sl)  [0]=upwind, [1]=downwind
    low_boundary_coefs   (bpc, bsl) or (2, bpc, bsl)
    high_boundary_coefs  (bpc[+offside], bsl) or (2, bpc+offside, bsl)
    """

    def __init__(self, T, axis, wind, derivative_order, approximation_order, dx, len_, stencil_length,
                 stencil_coefs, boundary_stencil_length, boundary_point_count, low_boundary_coefs,
                 high_boundary_coefs, offside, coefficients, coeff_func):
        self.T = np.dtype(T)
        self.axis = int(axis)                  # the {N} type parameter (1-based)
        self.wind = bool(wind)                 # the Wind type parameter
        self.derivative_order = derivative_order
        self.approximation_order = approximation_order
        self.dx = dx
        self.len = len_
        self.stencil_length = stencil_length
        self.stencil_coefs = stencil_coefs
        self.boundary_stencil_length = boundary_stencil_length
        self.boundary_point_count = boundary_point_count
        self.low_boundary_coefs = low_boundary_coefs
        self.high_boundary_coefs = high_boundary_coefs
        self.offside = offside
        self.coefficients = coefficients
        self.coeff_func = coeff_func
        self._version = 0                      # bumped by update_coefficients_ (plan cache key)

    @property
    def nonuniform(self):
        return np.ndim(self.dx) != 0

    @property
    def size(self):                            # abstract_operator_functions.jl:120-123
        return (self.len, self.len + 2)

    def _with_coefficients(self, coefficients):
        return DerivativeOperator(self.T, self.axis, self.wind, self.derivative_order, self.approximation_order,
                                  self.dx, self.len, self.stencil_length, self.stencil_coefs,
                                  self.boundary_stencil_length, self.boundary_point_count,
                                  self.low_boundary_coefs, self.high_boundary_coefs, self.offside,
                                  coefficients, self.coeff_func)

    def __rmul__(self, c):
        # c*A for a Number or a Vector: derivative_operator_functions.jl:165-197
        if isinstance(c, numbers.Number) or np.ndim(c) == 0:
            return self._with_coefficients((self.T.type(c) * self.coefficients).astype(self.T))
        c = np.asarray(c, dtype=self.T)
        if c.shape != (self.len,):
            raise ValueError(f"DimensionMismatch: length of c ({c.size}) must match length of A ({self.len})")
        return self._with_coefficients(c * self.coefficients)

    def __neg__(self):
        return (-1) * self

    def __mul__(self, other):
        from .bc import AbstractBC
        from .apply import mul_alloc
        if isinstance(other, AbstractBC):
            return GhostDerivativeOperator(self, other)       # ghost_derivative_operator.jl:7-9
        if isinstance(other, AbstractDiffEqLinearOperator):
            return _compose(self, other)
        return mul_alloc(self, other)                          # derivative_operator_functions.jl:150-163

    def update_coefficients_(self, u, p, t):
        # abstract_operator_functions.jl:190-194: coeff_func(coefficients, u, p, t) mutates in place
        if callable(self.coeff_func):
            try:
                self.coeff_func(self.coefficients, u, p, t)
            except TypeError:
                return
            self._version += 1

    def set_coefficients(self, c):
        self.coefficients[:] = np.asarray(c, dtype=self.T)
        self._version += 1


class CenteredDifference(metaclass=_Indexable):
    """CenteredDifference{N}(derivative_order, approximation_order, dx, len, coeff_func = 1)
    derivative_operator.jl:81-135 (dx::Real), :150-205 (dx::AbstractVector, len+1 spacings)."""

    def __new__(cls, derivative_order, approximation_order, dx, len_, coeff_func=1, *, axis=1, dtype=None):
        dt = _infer_dtype(dx, dtype)
        T = dt.type
        d, a, n = int(derivative_order), int(approximation_order), int(len_)
        sl = d + a - 1 + (d + a) % 2
        bsl = d + a
        bpc = sl // 2 - 1
        if np.ndim(dx) == 0:
            assert a > 1, "approximation_order must be greater than 1."
            dx = T(dx)
            with np.errstate(all="ignore"):
                scale = T(1) / dx ** d                                   # (1 / dx^derivative_order)
            stencil = scale * calculate_weights(d, 0, range(-(sl // 2), sl // 2 + 1), dt)
            low = np.zeros((bpc, bsl), dtype=dt)
            for i in range(1, bpc + 1):                                   # L_boundary_deriv_spots = 1..bpc
                low[i - 1] = scale * calculate_weights(d, i, range(bsl), dt)
            sgn = T(1) if d % 2 == 0 else T(-1)
            high = (low * sgn)[::-1, ::-1].copy()                         # reverse(map(reverse, low*(-1)^d))
        else:
            dx = np.ascontiguousarray(dx, dtype=dt)
            if dx.shape != (n + 1,):
                raise ValueError(f"dx must hold len+1 = {n + 1} grid spacings, got {dx.shape}")
            R = sl // 2
            nint = n - 2 * bpc
            stencil = np.zeros((max(nint, 0), sl), dtype=dt)
            for r in range(nint):                                         # interior_x = (bpc+2):(len+1-bpc)
                i = bpc + 2 + r
                sx = [T(0)] * sl
                for idx in range(1, R + 1):                               # generate_coordinates :137-148
                    sx[R + idx] = sx[R + idx - 1] + dx[i + idx - 2]
                    sx[R - idx] = sx[R - idx + 1] - dx[i - idx - 1]
                stencil[r] = calculate_weights(d, 0, sx, dt)
            lbx = np.concatenate([[T(0)], julia_cumsum(dx[:bsl - 1], dt)])
            hbx = julia_cumsum(dx[n + 1 - bsl:], dt)
            low = np.zeros((bpc, bsl), dtype=dt)
            high = np.zeros((bpc, bsl), dtype=dt)
            for i in range(1, bpc + 1):
                low[i - 1] = calculate_weights(d, lbx[i], lbx, dt)
            for j in range(bpc):                                          # for i in bpc:-1:1
                i = bpc - j
                high[j] = calculate_weights(d, hbx[bsl - 1 - i], hbx, dt)
        coefficients = compute_coeffs(coeff_func, n, dt)
        return DerivativeOperator(dt, axis, False, d, a, dx, n, sl, stencil, bsl, bpc, low, high, 0,
                                  coefficients, coeff_func)


class UpwindDifference(metaclass=_Indexable):
    """UpwindDifference{N}(derivative_order, approximation_order, dx, len, coeff_func = 1; offside = 0)
    derivative_operator.jl:454-512 (dx::Real), :515-639 (dx::AbstractVector).  The weights are
    computed in Float64 and converted to T, as the reference does (0.0 literals, Float64 ranges)."""

    def __new__(cls, derivative_order, approximation_order, dx, len_, coeff_func=1, *, offside=0, axis=1, dtype=None):
        dt = _infer_dtype(dx, dtype)
        T = dt.type
        f64 = np.float64
        d, a, n, off = int(derivative_order), int(approximation_order), int(len_), int(offside)
        assert off > -1, "Number of offside points should be non-negative"
        assert off <= (d + a - 1) // 2, "Number of offside points should not exceed the primary wind points"
        sl = bsl = d + a
        bpc = bsl - 2 - off
        nhigh = bpc + off
        if np.ndim(dx) == 0:
            dx = T(dx)
            with np.errstate(all="ignore"):
                scale = f64(T(1) / dx ** d)                               # (1/dx^d)::T promoted to Float64
                hscale = f64((T(-1) / dx) ** d)                           # ((-1/dx)^d)::T
            stencil = (scale * calculate_weights(d, 0.0, [f64(k - off) for k in range(sl)], f64)).astype(dt)
            low = np.zeros((bpc, bsl), dtype=dt)
            for i in range(1, bpc + 1):
                low[i - 1] = (scale * calculate_weights(d, f64(i), range(bsl), f64)).astype(dt)
            high = np.zeros((nhigh, bsl), dtype=dt)
            for i in range(1, nhigh + 1):
                high[i - 1] = (hscale * calculate_weights(d, f64(-i), [f64(-k) for k in range(bsl)], f64)).astype(dt)
        else:
            dx = np.ascontiguousarray(dx, dtype=dt)
            if dx.shape != (n + 1,):
                raise ValueError(f"dx must hold len+1 = {n + 1} grid spacings, got {dx.shape}")
            X = np.concatenate([[0.0], julia_cumsum(dx, dt).astype(f64)])   # x = [0.0, cumsum(dx)...]
            nint = n - 2 * bpc

            def W(x0, start, m):
                return calculate_weights(d, x0, X[start:start + m], f64).astype(dt)

            low = np.zeros((2, bpc, bsl), dtype=dt)
            for i in range(1, bpc + 1):
                if off == 0:
                    low[0, i - 1] = W(X[i], i, bsl)
                elif i <= off - 1:
                    low[0, i - 1] = W(X[i], 0, bsl)
                else:
                    low[0, i - 1] = W(X[i], i - off, bsl)
                low[1, i - 1] = W(X[i], 0, bsl)
            stencil = np.zeros((2, max(nint, 0), sl), dtype=dt)
            for r in range(nint):
                i = bpc + 1 + r
                stencil[0, r] = W(X[i], i - off, sl)
                stencil[1, r] = W(X[i], i - sl + 1 + off, sl)
            high = np.zeros((2, nhigh, bsl), dtype=dt)
            for r in range(nhigh):
                i = n - bpc + 1 - off + r
                high[0, r] = W(X[i], n - bsl + 2, bsl)
            if off == 0:
                for r in range(bpc):
                    i = n - bpc + 1 + r
                    high[1, r] = W(X[i], i - sl + 1, bsl)
            else:
                r = 0
                for i in range(n - bpc + 1 - off, n - off + 2):
                    high[1, r] = W(X[i], i - sl + 1 + off, bsl)
                    r += 1
                if off >= 2:
                    for i in range(n - off + 2, n + 1):
                        high[1, r] = W(X[i], n - sl + 2, bsl)
                        r += 1
        coefficients = compute_coeffs(coeff_func, n, dt)
        return DerivativeOperator(dt, axis, True, d, a, dx, n, sl, stencil, bsl, bpc, low, high, off,
                                  coefficients, coeff_func)


def Laplacian(aor, dxyz, s, coeff_func=1):
    """derivative_operator.jl:765-768."""
    ops = [CenteredDifference[i + 1](2, aor, dxyz[i], s[i], coeff_func) for i in range(len(s))]
    return DiffEqOperatorCombination(tuple(ops)) if len(ops) > 1 else ops[0]


class GhostDerivativeOperator(AbstractDiffEqLinearOperator):
    """L*Q (ghost_derivative_operator.jl:1-9)."""

    def __init__(self, L, Q):
        if L.T != Q.T:
            raise TypeError("operator and boundary condition element types differ")
        self.L, self.Q = L, Q

    @property
    def T(self):
        return self.L.T

    @property
    def size(self):                                # ghost_derivative_operator.jl:79-82
        return (self.L.len, self.L.len)

    def __rmul__(self, c):                         # :66-76  (c*A.L)*A.Q
        return GhostDerivativeOperator(c * self.L, self.Q)

    def __neg__(self):
        return (-1) * self

    def __mul__(self, u):
        from .apply import mul_alloc
        if isinstance(u, AbstractDiffEqLinearOperator):
            return _compose(self, u)
        return mul_alloc(self, u)                  # :26-37

    def update_coefficients_(self, u, p, t):       # :61-63
        self.L.update_coefficients_(u, p, t)


class DiffEqOperatorCombination(AbstractDiffEqLinearOperator):
    """composite_operators.jl:15-83 (`+` of operators)."""

    def __init__(self, ops):
        ops = tuple(ops)
        assert len(ops) >= 1
        # composite_operators.jl:22-24 asserts equal sizes; operators along different axes of a
        # non-cubic array necessarily differ, so only same-axis operators are held to it here
        for i, op in enumerate(ops[1:], 2):
            for prev in ops[:i - 1]:
                if getattr(_base(op), "axis", None) == getattr(_base(prev), "axis", None):
                    assert op.size == prev.size, f"Operators must be of the same size to be combined! Mismatch at operator {i}"
        self.ops = ops

    @property
    def T(self):
        return self.ops[0].T

    @property
    def size(self):
        return self.ops[0].size

    def __rmul__(self, c):
        return DiffEqOperatorCombination(tuple(c * op for op in self.ops))

    def __neg__(self):
        return (-1) * self

    def __mul__(self, other):
        from .bc import AbstractBC
        from .apply import mul_alloc
        if isinstance(other, AbstractBC):
            # (sum L)*Q = sum(L*Q)   ghost_derivative_operator.jl:11-13
            return DiffEqOperatorCombination(tuple(op * other for op in self.ops))
        if isinstance(other, AbstractDiffEqLinearOperator):
            return _compose(self, other)
        return mul_alloc(self, other)              # composite_operators.jl:64-65

    def update_coefficients_(self, u, p, t):
        for op in self.ops:
            op.update_coefficients_(u, p, t)
