"""Host-side mirror of the reference's boundary-condition operators (constructors only; ghost
values are evaluated inside the CUDA kernels, never materialised):

  bc_operators.jl:17-19     PeriodicBC
  bc_operators.jl:21-61     RobinBC (dx::T and dx::Vector)
  bc_operators.jl:85-170    GeneralBC
  bc_operators.jl:173-183   NeumannBC, DirichletBC, Dirichlet0BC, Neumann0BC
  multi_dim_bc_operators.jl:54-61,93-163   MultiDimDirectionalBC / ComposedMultiDimBC / MultiDimBC{dim}
  multi_dim_bc_operators.jl:178-194        compose
  boundary_padded_arrays.jl                 BoundaryPadded* (here: a lazy (Q, u) pair)

Every affine BC reduces to the 4-tuple (a_l, b_l, a_r, b_r): ghost_l = a_l . u[1:K] + b_l,
ghost_r = a_r . u[end-K+1:end] + b_r   (bc_operators.jl:188-191).
"""
from __future__ import annotations

import numpy as np

from .operators import _Indexable, calculate_weights


def _dt(*vals, dtype=None):
    if dtype is not None:
        return np.dtype(dtype)
    for v in vals:
        if isinstance(v, (np.ndarray, np.generic)) and np.asarray(v).dtype == np.float32:
            return np.dtype(np.float32)
        if isinstance(v, (tuple, list)) and any(isinstance(e, np.float32) for e in v):
            return np.dtype(np.float32)
    return np.dtype(np.float64)


class AbstractBC:
    T: np.dtype

    def __rmul__(self, L):
        return NotImplemented


class AtomicBC(AbstractBC):
    def __mul__(self, u):
        from .device import DeviceArray
        nd = u.ndim if isinstance(u, (np.ndarray, DeviceArray)) else np.ndim(u)
        if nd == 1:
            return BoundaryPadded(self, u)                        # bc_operators.jl:188-192
        # Base.:*(BC::AtomicBC, u::AbstractArray) = MultiDimBC{1}(BC, size(u)) * u   (multi_dim_bc_operators.jl:210)
        return MultiDimBC[1](self, u.shape) * u


class AffineBC(AtomicBC):
    """a_l, b_l, a_r, b_r  (bc_operators.jl:21-25, :85-89)."""

    def __init__(self, a_l, b_l, a_r, b_r, T):
        self.T = np.dtype(T)
        self.a_l = np.ascontiguousarray(a_l, dtype=T)
        self.a_r = np.ascontiguousarray(a_r, dtype=T)
        self.b_l = self.T.type(b_l)
        self.b_r = self.T.type(b_r)


class PeriodicBC(AtomicBC, metaclass=_Indexable):
    def __new__(cls, T=np.float64, s=None, *, axis=None):
        if s is not None or axis is not None:
            # PeriodicBC{dim}(T, s) / PeriodicBC(T, s)  (multi_dim_bc_operators.jl:113-114)
            q = object.__new__(cls)
            q.T = np.dtype(T)
            return MultiDimBC[axis](q, s) if axis is not None else MultiDimBC(q, s)
        q = object.__new__(cls)
        q.T = np.dtype(T)
        return q


def _robin(l, r, dx, order, T):
    """RobinBC(l, r, dx, order): bc_operators.jl:26-39 (dx::T), :40-60 (dx::AbstractVector)."""
    dt = np.dtype(T)
    t = dt.type
    al, bl, cl = (t(v) for v in l)
    ar, br, cr = (t(v) for v in r)
    order = int(order)
    s = calculate_weights(1, 1, [k + 1 for k in range(order + 1)], dt)   # weights(1, one(T), 1:order+1)
    with np.errstate(all="ignore"):                                      # Dirichlet relies on 1/0 = Inf (:176-178)
        if np.ndim(dx) == 0:
            dx = t(dx)
            a_l = -s[1:] / (al * dx / bl + s[0])
            a_r = s[:0:-1] / (ar * dx / br - s[0])
            b_l = cl / (al + bl * s[0] / dx)
            b_r = cr / (ar - br * s[0] / dx)
        else:
            dx = np.asarray(dx, dtype=dt)
            m = order + 1
            dx_l, dx_r = dx[:m], dx[len(dx) - m:]
            denom_l = al + bl * s[0] / dx_l[0]
            denom_r = ar - br * s[0] / dx_r[-1]
            a_l = -bl * s[1:] / (denom_l * dx_l[1:])
            a_r = br * s[:0:-1] / (denom_r * dx_r[:-1])
            b_l = cl / denom_l
            b_r = cr / denom_r
    return AffineBC(a_l, b_l, a_r, b_r, dt)


def _is_size(x):
    return isinstance(x, (tuple, list)) and len(x) > 0 and all(isinstance(v, (int, np.integer)) for v in x)


class RobinBC(metaclass=_Indexable):
    """RobinBC((αl,βl,γl), (αr,βr,γr), dx, order=1)                      -> atomic BC
       RobinBC{dim}(l, r, dx, order, size(u))                              -> MultiDimDirectionalBC
       RobinBC(l, r, (dx,dy,dz), order, size(u))                           -> tuple, one per dim
       (multi_dim_bc_operators.jl:138-149)"""

    def __new__(cls, l, r, dx, order=1, s=None, *, axis=None, dtype=None):
        if s is None and axis is None:
            return _robin(l, r, dx, order, _dt(dx, l, r, dtype=dtype))
        if axis is not None:
            return MultiDimBC[axis](_robin(l, r, dx, order, _dt(dx, l, r, dtype=dtype)), s)
        return tuple(MultiDimBC[d + 1](_robin(l, r, dx[d], order, _dt(dx[d], l, r, dtype=dtype)), s) for d in range(len(s)))


class GeneralBC(metaclass=_Indexable):
    """GeneralBC(αl, αr, dx, order=1): α[1] + α[2]u + α[3]u' + ... = 0   (bc_operators.jl:85-170)."""

    def __new__(cls, alpha_l, alpha_r, dx, order=1, s=None, *, axis=None, dtype=None):
        if s is None and axis is None:
            return _general(alpha_l, alpha_r, dx, order, _dt(dx, dtype=dtype))
        if axis is not None:
            return MultiDimBC[axis](_general(alpha_l, alpha_r, dx, order, _dt(dx, dtype=dtype)), s)
        return tuple(MultiDimBC[d + 1](_general(alpha_l, alpha_r, dx[d], order, _dt(dx[d], dtype=dtype)), s) for d in range(len(s)))


def _general(alpha_l, alpha_r, dx, order, T):
    dt = np.dtype(T)
    t = dt.type
    order = int(order)
    al = [t(v) for v in alpha_l]
    ar = [t(v) for v in alpha_r]
    nl, nr = len(al), len(ar)
    cl, cr = order + nl - 2, order + nr - 2
    vec = np.ndim(dx) != 0
    if vec:
        dxa = np.asarray(dx, dtype=dt)
        dx_l = dxa[:cl]
        dx_r = dxa[len(dxa) - cr:][::-1]
    S_l = np.zeros((nl - 2, cl), dtype=dt)
    S_r = np.zeros((nr - 2, cr), dtype=dt)
    for i in range(1, nl - 1):
        w = calculate_weights(i, 1, [k + 1 for k in range(order + i)], dt)
        row = np.concatenate([w, np.zeros(nl - 2 - i, dtype=dt)])
        S_l[i - 1] = row / (dx_l ** i if vec else t(dx) ** i)
    for i in range(1, nr - 1):
        w = calculate_weights(i, order + i, [k + 1 for k in range(order + i)], dt)
        row = np.concatenate([w, np.zeros(nr - 2 - i, dtype=dt)])
        S_r[i - 1] = row / (dx_r ** i if vec else t(dx) ** i)
    s0_l, Sl = S_l[:, 0], S_l[:, 1:]
    s0_r, Sr = S_r[:, -1], S_r[:, -2::-1]
    dotl = t(0)
    for i in range(nl - 2):
        dotl = dotl + al[2 + i] * s0_l[i]
    dotr = t(0)
    for i in range(nr - 2):
        dotr = dotr + ar[2 + i] * s0_r[i]
    denoml = al[1] + dotl
    denomr = ar[1] + dotr
    a_l = np.zeros(cl - 1, dtype=dt)
    for j in range(cl - 1):
        acc = t(0)
        for i in range(nl - 2):
            acc = acc + al[2 + i] * Sl[i, j]
        a_l[j] = -acc / denoml
    a_r = np.zeros(cr - 1, dtype=dt)
    for j in range(cr - 1):
        acc = t(0)
        for i in range(nr - 2):
            acc = acc + ar[2 + i] * Sr[i, j]
        a_r[cr - 2 - j] = -acc / denomr
    return AffineBC(a_l, -al[0] / denoml, a_r, -ar[0] / denomr, dt)


class NeumannBC(metaclass=_Indexable):
    """NeumannBC((αl, αr), dx, order=1) = RobinBC((0,1,αl), (0,1,αr), dx, order)  (bc_operators.jl:173-175)."""

    def __new__(cls, alpha, dx, order=1, s=None, *, axis=None, dtype=None):
        return RobinBC((0, 1, alpha[0]), (0, 1, alpha[1]), dx, order, s, axis=axis, dtype=_dt(dx, alpha, dtype=dtype))


class DirichletBC(metaclass=_Indexable):
    """DirichletBC(αl, αr) = RobinBC((1,0,αl), (1,0,αr), 1, 2)  (bc_operators.jl:176-178);
    DirichletBC{dim}(αl, αr, s) / DirichletBC(αl, αr, s)  (multi_dim_bc_operators.jl:123-128)."""

    def __new__(cls, al, ar, s=None, *, axis=None, dtype=None):
        dt = _dt(al, ar, dtype=dtype)
        if s is None and axis is None:
            return _robin((1, 0, al), (1, 0, ar), 1, 2, dt)
        if axis is not None:
            return MultiDimBC[axis](_robin((1, 0, al), (1, 0, ar), 1, 2, dt), s)
        return tuple(MultiDimBC[d + 1](_robin((1, 0, al), (1, 0, ar), 1, 2, dt), s) for d in range(len(s)))


class Dirichlet0BC(metaclass=_Indexable):
    def __new__(cls, T=np.float64, s=None, *, axis=None):
        return DirichletBC(0, 0, s, axis=axis, dtype=T)


class Neumann0BC(metaclass=_Indexable):
    """Neumann0BC(dx, order=1)  (bc_operators.jl:181-183);  Neumann0BC(T, dxyz, order, s)  (multi_dim :133-136)."""

    def __new__(cls, *args, axis=None, dtype=None):
        if len(args) >= 1 and isinstance(args[0], (type, np.dtype)):
            T, dx, order, s = args
            return NeumannBC((0, 0), dx, order, s, axis=axis, dtype=T)
        dx = args[0]
        order = args[1] if len(args) > 1 else 1
        return NeumannBC((0, 0), dx, order, dtype=_dt(dx, dtype=dtype))


# ---- multi-dimensional BCs ----------------------------------------------------------------------
class MultiDimDirectionalBC(AbstractBC):
    """An (N-1)-dimensional array of atomic BCs extending dimension `axis` (1-based)
    (multi_dim_bc_operators.jl:54-57).  `BCs` is either one atomic BC (the `fill(BC, ...)` case,
    :97-100) or a numpy object array with one BC per boundary pencil."""

    def __init__(self, BCs, axis, ndims, face_shape):
        self.BCs = BCs
        self.axis = int(axis)
        self.ndims = int(ndims)
        self.face_shape = tuple(int(v) for v in face_shape)
        first = BCs if isinstance(BCs, AtomicBC) else BCs.flat[0]
        self.T = first.T

    @property
    def uniform(self):
        return isinstance(self.BCs, AtomicBC)

    def __mul__(self, u):
        assert tuple(np.delete(np.array(u.shape), self.axis - 1)) == self.face_shape, \
            f"Size of the BCs array in the MultiDimBC is incorrect, needs to be {tuple(np.delete(np.array(u.shape), self.axis - 1))} to extend dimension {self.axis}, got {self.face_shape}"
        return BoundaryPadded(self, u)                             # :212-219


class ComposedMultiDimBC(AbstractBC):
    """multi_dim_bc_operators.jl:59-61; BCs[d] extends dimension d+1."""

    def __init__(self, parts):
        self.parts = list(parts)
        self.T = parts[0].T
        self.ndims = parts[0].ndims

    def __mul__(self, u):
        for d, part in enumerate(self.parts):                      # :232-234
            assert tuple(np.delete(np.array(u.shape), d)) == part.face_shape, \
                f"Size of the BCs array for dimension {d + 1} in the MultiDimBC is incorrect"
        return BoundaryPadded(self, u)                             # :230-242


class _MultiDimBCMeta(type):
    def __getitem__(cls, dim):
        dim = int(dim)

        def make(BC, s=None):
            if isinstance(BC, np.ndarray):                         # MultiDimBC{dim}(BC::Array)  :93-95
                return MultiDimDirectionalBC(BC, dim, BC.ndim + 1, BC.shape)
            face = tuple(v for i, v in enumerate(s) if i != dim - 1)   # MultiDimBC{dim}(BC, s)  :97-100
            return MultiDimDirectionalBC(BC, dim, len(s), face)
        return make


class MultiDimBC(metaclass=_MultiDimBCMeta):
    """MultiDimBC{dim}(BC, size(u)); MultiDimBC(BC, size(u)) -> tuple for all dims (:105-110)."""

    def __new__(cls, BC, s):
        return tuple(MultiDimBC[d + 1](BC, s) for d in range(len(s)))


def compose(*BCs):
    """Q = compose(Qx, Qy, Qz)  (multi_dim_bc_operators.jl:178-192)."""
    N = BCs[0].ndims
    Ds = [b.axis for b in BCs]
    if len(BCs) != N:
        raise ValueError("There must be enough BCs to cover every dimension - check that the number of MultiDimBCs == N")
    if len(set(Ds)) != N:
        raise ValueError("There are multiple boundary conditions that extend along the same dimension - make sure every dimension has a unique extension")
    return ComposedMultiDimBC([b for _, b in sorted(zip(Ds, BCs), key=lambda t: t[0])])


class BoundaryPadded:
    """Lazy result of Q*u (BoundaryPaddedVector / BoundaryPaddedArray / ComposedBoundaryPaddedArray,
    boundary_padded_arrays.jl): keeps (Q, u); the ghosts are evaluated in-kernel when an operator is
    applied, or on the host by `to_array()` (= Array(Q*u), corners 0, :158-180) for inspection."""

    def __init__(self, Q, u):
        self.Q, self.u = Q, u

    @property
    def padded_axes(self):
        if isinstance(self.Q, ComposedMultiDimBC):
            return list(range(1, self.Q.ndims + 1))
        if isinstance(self.Q, MultiDimDirectionalBC):
            return [self.Q.axis]
        return [1]

    def to_array(self):
        from .apply import materialize_padded
        return materialize_padded(self)
