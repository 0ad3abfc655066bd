"""Vector-calculus operators, operator composition and `nonlinear_diffusion!` on top of the fused application path
(SURVEY 8f-3 / 8f-4): host-side mirrors of

  derivative_operators/vector_calculus_operators.jl:5-88     Gradient / Divergence / Curl (+ `c *`)
  derivative_operators/vector_calculus_convolutions.jl:3-87  their mul! on pre-padded arrays
  derivative_operators/derivative_operator.jl:31-77          nonlinear_diffusion! / nonlinear_diffusion
  composite_operators.jl:86-212                              DiffEqOperatorComposition (`∘`, `*`, mul! through caches)

Every derivative is one libdeo_b200 plan on a DeviceArray (or a component window of one); nothing is computed on the host.
Semantics on pre-padded input are the reference's generic N-D path (vector_calculus_convolutions.jl:30-37, :66-73): the
differentiated axis reads its own ghost layer, the other padded axes are stripped with `2:end-1`.  (The reference's
specialised 2-D/3-D Curl loops index the perpendicular axes without that shift, vector_calculus_convolutions.jl:878-925 --
indistinguishable on its own separable test fields, curl_operator.jl:9-24; the intended semantics are implemented here,
the same decision SURVEY 2.1-2 records for the composite conv path.)
"""
from __future__ import annotations

from math import comb

import numpy as np

from . import _lib
from .apply import mul_
from .device import DeviceArray
from .operators import AbstractDiffEqLinearOperator, CenteredDifference


def _dev(a):
    return a if isinstance(a, DeviceArray) else DeviceArray.from_host(np.asfortranarray(a))


def _like_input(out: DeviceArray, like):
    return out if isinstance(like, DeviceArray) else out.to_host()


class _AxisOps:
    def __init__(self, ops):
        self.ops = list(ops)
        self.T = self.ops[0].T
        self.N = len(self.ops)

    def __rmul__(self, c):                               # vector_calculus_operators.jl:75-88
        return type(self)._from_ops([c * op for op in self.ops])

    @classmethod
    def _from_ops(cls, ops):
        o = cls.__new__(cls)
        _AxisOps.__init__(o, ops)
        return o

    def _lens(self):
        return tuple(op.len for op in self.ops)


class GradientOperator(_AxisOps):
    """Gradient(approximation_order, dx, len; coeff_func=1)  (vector_calculus_operators.jl:9-17)."""

    def __init__(self, approximation_order, dx, len_, coeff_func=1, dtype=None):
        N = len(len_)
        super().__init__([CenteredDifference[n + 1](1, approximation_order, dx[n], len_[n], coeff_func, dtype=dtype) for n in range(N)])

    def __mul__(self, M):                                # :47-54 -> out[.., n] = d/dx_n M
        Md = _dev(M)
        lens = self._lens()
        assert Md.shape == tuple(l + 2 for l in lens), f"Gradient: size(M) = {Md.shape} must be len .+ 2 = {tuple(l + 2 for l in lens)}"
        out = DeviceArray(lens + (self.N,), self.T)
        npts = int(np.prod(lens))
        for n, op in enumerate(self.ops):
            mul_(out.view(n * npts, lens), op, Md)
        return _like_input(out, M)


class DivergenceOperator(_AxisOps):
    """Divergence(approximation_order, dx, len; coeff_func=1)  (vector_calculus_operators.jl:38-45)."""

    def __init__(self, approximation_order, dx, len_, coeff_func=1, dtype=None):
        N = len(len_)
        super().__init__([CenteredDifference[n + 1](1, approximation_order, dx[n], len_[n], coeff_func, dtype=dtype) for n in range(N)])

    def __mul__(self, M):                                # :65-73 -> out = sum_n d/dx_n M[.., n]
        Md = _dev(M)
        lens = self._lens()
        pshape = tuple(l + 2 for l in lens)
        assert Md.shape == pshape + (self.N,), f"Divergence: size(M) = {Md.shape} must be (len .+ 2..., {self.N})"
        out = DeviceArray(lens, self.T)
        npad = int(np.prod(pshape))
        for n, op in enumerate(self.ops):
            mul_(out, op, Md.view(n * npad, pshape), overwrite=(n == 0))
        return _like_input(out, M)


class CurlOperator(_AxisOps):
    """Curl(approximation_order, dx, len; coeff_func=1) for 3-D vectors stored as (n1, n2, n3, 3) space-tensors
    (vector_calculus_operators.jl:21-34)."""

    def __init__(self, approximation_order, dx, len_, coeff_func=1, dtype=None):
        assert len(len_) == 3 and len_[0] == len_[1] == len_[2], "All dimensions must have equal no. of grid points"
        super().__init__([CenteredDifference[n + 1](1, approximation_order, dx[n], len_[n], coeff_func, dtype=dtype) for n in range(3)])

    def __mul__(self, M):                                # :56-63, vector_calculus_convolutions.jl:844-925
        Md = _dev(M)
        lens = self._lens()
        pshape = tuple(l + 2 for l in lens)
        assert Md.shape == pshape + (3,), f"Curl: size(M) = {Md.shape} must be (len .+ 2..., 3)"
        out = DeviceArray(lens + (3,), self.T)
        npts, npad = int(np.prod(lens)), int(np.prod(pshape))
        comp_in = [Md.view(c * npad, pshape) for c in range(3)]
        comp_out = [out.view(c * npts, lens) for c in range(3)]
        if not hasattr(self, "_neg"):
            self._neg = [(-1) * op for op in self.ops]       # -c*w is exact; kept so that their plans are cached
        D, Dn = self.ops, self._neg
        # (curl u)_1 = d2 u3 - d3 u2 ; (curl u)_2 = d3 u1 - d1 u3 ; (curl u)_3 = d1 u2 - d2 u1
        for c, (plus, u_plus, minus, u_minus) in enumerate(((1, 2, 2, 1), (2, 0, 0, 2), (0, 1, 1, 0))):
            mul_(comp_out[c], D[plus], comp_in[u_plus])
            mul_(comp_out[c], Dn[minus], comp_in[u_minus], overwrite=False)
        return _like_input(out, M)


def Gradient(approximation_order, dx, len_, coeff_func=1, dtype=None):
    return GradientOperator(approximation_order, dx, len_, coeff_func, dtype)


def Divergence(approximation_order, dx, len_, coeff_func=1, dtype=None):
    return DivergenceOperator(approximation_order, dx, len_, coeff_func, dtype)


def Curl(approximation_order, dx, len_, coeff_func=1, dtype=None):
    return CurlOperator(approximation_order, dx, len_, coeff_func, dtype)


# ---- nonlinear_diffusion! (derivative_operator.jl:31-77) --------------------------------------------------------------------
def _muladd(dst: DeviceArray, a: DeviceArray, b: DeviceArray, scale, accumulate):
    _lib.check(_lib.load().deo_buffer_muladd(dst._h, a._h, b._h, float(scale), int(bool(accumulate)), dst.size, _lib.dtype_code(dst.dtype)))


def nonlinear_diffusion_(du, second_differential_order, first_differential_order, approx_order, p, q, dx, nknots, axis=1):
    """du = D^f q .* D^s p + sum_l C(s,l) D^(l+f) q .* D^(s-l) p + D^(f+s) q .* p[2:end-1]   (1-D; p, q carry their ghosts:
    q = bc*u, p = a function of q).  Every derivative is a plan on the device, every product one elementwise kernel."""
    assert approx_order > 1, "approximation_order must be greater than 1."
    pd, qd = _dev(p), _dev(q)
    T = qd.dtype
    assert pd.shape == qd.shape == (nknots + 2,), "p and q must be padded vectors of length nknots + 2"
    out = du if isinstance(du, DeviceArray) else DeviceArray((nknots,), T)
    ta, tb = DeviceArray((nknots,), T), DeviceArray((nknots,), T)
    s, f = second_differential_order, first_differential_order

    def D(order, v, dst):
        return mul_(dst, CenteredDifference[axis](order, approx_order, dx, nknots, dtype=T), v)

    inner = lambda v: v.view(1, (nknots,))                   # v[2:nknots+1]
    if f > 0:
        _muladd(out, D(f, qd, ta), D(s, pd, tb), 1.0, False)
    else:
        _muladd(out, inner(qd), D(s, pd, tb), 1.0, False)
    for l in range(1, s):
        _muladd(out, D(l + f, qd, ta), D(s - l, pd, tb), comb(s, l), True)
    _muladd(out, D(f + s, qd, ta), inner(pd), 1.0, True)
    if not isinstance(du, DeviceArray):
        du[...] = out.to_host()
        return du
    return out


def nonlinear_diffusion(second_differential_order, first_differential_order, approx_order, p, q, dx, nknots, axis=1):
    T = np.asarray(q).dtype if not isinstance(q, DeviceArray) else q.dtype
    out = nonlinear_diffusion_(DeviceArray((nknots,), T), second_differential_order, first_differential_order, approx_order, p, q, dx, nknots, axis)
    return _like_input(out, q)


# ---- DiffEqOperatorComposition (composite_operators.jl:86-212) ----------------------------------------------------------------
class DiffEqOperatorComposition(AbstractDiffEqLinearOperator):
    """Operators stored in the order of application; `*` folds them over the input, `mul_` runs them through device caches
    (composite_operators.jl:150-152, :186-192).  Each member is applied by its own fused plan."""

    def __init__(self, ops):
        self.ops = tuple(ops)
        assert len(self.ops) >= 1
        self._caches = {}

    @property
    def T(self):
        return self.ops[0].T

    def __mul__(self, x):
        if isinstance(x, AbstractDiffEqLinearOperator):     # L1 * L2 = apply L2 first (:116-133)
            other = x.ops if isinstance(x, DiffEqOperatorComposition) else (x,)
            return DiffEqOperatorComposition(tuple(other) + self.ops)
        acc = x
        for op in self.ops:                                  # foldl((acc, op) -> op * acc, L.ops; init = x)
            acc = op * acc
        return acc

    def mul_(self, y, b):
        """mul!(y, L, b) through per-shape device caches (composite_operators.jl:186-192); DeviceArray operands."""
        acc = b
        for i, op in enumerate(self.ops[:-1]):
            key = (i, acc.shape)
            if key not in self._caches:
                self._caches[key] = op * acc                 # allocates the cache with the right shape on first use
            else:
                mul_(self._caches[key], op, acc)
            acc = self._caches[key]
        return mul_(y, self.ops[-1], acc)


# ---- concretization of L*Q and `\` (ghost_derivative_operator.jl:39-58, concretization.jl:914-958) -----------------------------
def concretize(G, shape, dtype=None, batch=256):
    """Array(A::GhostDerivativeOperator, s) -> (A_l, A_b): the linear part (prod(s) x prod(s), column-major unknowns) and
    the affine part of `u -> (L*Q) u`, obtained ON THE DEVICE by applying the fused plan to zero and to the unit vectors
    (`A_b = G*0`, `A_l e_k = G*e_k - A_b`); only the columns come back to the host.  Meant for the sizes the reference
    concretizes (tests, direct solves of small problems): cost is prod(s) applications."""
    shape = tuple(int(v) for v in np.atleast_1d(shape))
    T = np.dtype(dtype) if dtype is not None else np.dtype(G.T)
    n = int(np.prod(shape))
    e = DeviceArray.from_host(np.zeros(shape, dtype=T, order="F"))
    out = DeviceArray(shape, T)
    mul_(out, G, e)
    A_b = out.to_host().reshape(-1, order="F").astype(T)
    A_l = np.empty((n, n), dtype=T, order="F")
    host = np.zeros(n, dtype=T)
    for k in range(n):
        host[k] = 1
        e.upload(host.reshape(shape, order="F"))
        host[k] = 0
        mul_(out, G, e)
        A_l[:, k] = out.to_host().reshape(-1, order="F") - A_b
    return A_l, A_b


def ldiv(G, u):
    r"""A \ u for A = L*Q: solve A_l x = u - A_b with the concretized operator (ghost_derivative_operator.jl:39-58; the
    reference factorises the sparse matrix on the CPU, here the matrix comes from the device and LAPACK solves it)."""
    arr = u.to_host() if isinstance(u, DeviceArray) else np.asarray(u)
    A_l, A_b = concretize(G, arr.shape, arr.dtype)
    x = np.linalg.solve(A_l.astype(np.float64), arr.reshape(-1, order="F").astype(np.float64) - A_b)
    x = np.asfortranarray(x.astype(arr.dtype).reshape(arr.shape, order="F"))
    return DeviceArray.from_host(x) if isinstance(u, DeviceArray) else x


def compose_operators(L1, L2):
    """L1 ∘ L2 (composite_operators.jl:113-115): apply L2, then L1."""
    a = L2.ops if isinstance(L2, DiffEqOperatorComposition) else (L2,)
    b = L1.ops if isinstance(L1, DiffEqOperatorComposition) else (L1,)
    return DiffEqOperatorComposition(tuple(a) + tuple(b))
