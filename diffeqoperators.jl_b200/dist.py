"""Slab-decomposed application: one process per GPU, the last axis split into contiguous slabs, halo
planes exchanged by the library (NCCL send/recv over NVLink) overlapped with the interior planes.

Host-side responsibilities (this file): moving the 128-byte NCCL unique id between the ranks with
whatever process group the host runtime has (torch.distributed here; MPI / Distributed.jl in a Julia
deployment), slab bookkeeping, scatter/gather of host arrays.  No arithmetic happens here."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .apply import Plan, _bc_for_axis, _terms
from .bc import AffineBC
from .device import DeviceArray


def slab_bounds(n_last: int, nranks: int, rank: int):
    """[start, start+count) of `rank`'s planes along the last axis (same rule as deo_dist_slab;
    the library call is the authority, this wrapper only marshals)."""
    s = C.c_int64(0)
    c = C.c_int64(0)
    _lib.check(_lib.load().deo_dist_slab(int(n_last), int(nranks), int(rank), C.byref(s), C.byref(c)))
    return int(s.value), int(c.value)


def extended_slab(u_global: np.ndarray, rank: int, nranks: int, halo: int, fill=np.nan):
    """Host helper (tests, scatter): the [halo | own planes | halo] block of `rank`, halos filled from the
    neighbouring slabs where they exist and with `fill` at the physical faces (never read there)."""
    n = u_global.shape[-1]
    s, c = slab_bounds(n, nranks, rank)
    ext = np.full(u_global.shape[:-1] + (c + 2 * halo,), fill, dtype=u_global.dtype, order="F")
    lo, hi = max(s - halo, 0), min(s + c + halo, n)
    ext[..., lo - (s - halo):hi - (s - halo)] = u_global[..., lo:hi]
    return ext


class SlabContext:
    """Owns the library's NCCL communicator for this rank."""

    def __init__(self, rank: int, nranks: int, id_bytes: bytes):
        self.rank, self.nranks = int(rank), int(nranks)
        h = C.c_void_p()
        buf = C.create_string_buffer(bytes(id_bytes), _lib.DEO_DIST_ID_BYTES)
        _lib.check(_lib.load().deo_dist_init(buf, self.rank, self.nranks, C.byref(h)))
        self._h = h

    @staticmethod
    def new_unique_id() -> bytes:
        buf = C.create_string_buffer(_lib.DEO_DIST_ID_BYTES)
        _lib.check(_lib.load().deo_dist_unique_id(buf))
        return buf.raw

    @classmethod
    def from_torch_distributed(cls, device: int | None = None):
        """Bootstraps over an initialised torch.distributed process group (any backend): rank 0 creates
        the NCCL unique id and broadcasts its bytes."""
        import torch
        import torch.distributed as dist
        rank, world = dist.get_rank(), dist.get_world_size()
        if device is not None:
            _lib.check(_lib.load().deo_init(int(device)))
        backend = dist.get_backend()
        dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
        t = torch.zeros(_lib.DEO_DIST_ID_BYTES, dtype=torch.uint8, device=dev)
        if rank == 0:
            t.copy_(torch.frombuffer(bytearray(cls.new_unique_id()), dtype=torch.uint8))
        dist.broadcast(t, src=0)
        return cls(rank, world, bytes(t.cpu().numpy().tobytes()))

    def __del__(self):
        try:
            if self._h is not None:
                _lib.load().deo_dist_destroy(self._h)
                self._h = None
        except Exception:
            pass


class SlabPlan(Plan):
    """A fused operator application on this rank's slab of a global 3-D problem."""

    def __init__(self, A, global_shape, dtype, ctx: SlabContext | None = None, local_rank=None, flags=0):
        terms = _terms(A)
        nd = len(global_shape)
        bcs = []
        for ax in range(1, nd + 1):
            per_op = [_bc_for_axis(Q, ax, nd) for L, Q in terms if L.axis == ax]
            if not per_op:
                bcs.append(None)
                continue
            if any(b is None for b in per_op):
                raise AssertionError("slab plans need a boundary condition on every differentiated axis")
            # every operator of an axis must see the same ghosts (same rule as apply.build_plans): a sum such as
            # L1*Q1 + L2*Q2 along one axis cannot be fused into one pass
            first = per_op[0]
            for b in per_op[1:]:
                same = b is first or (isinstance(b, AffineBC) and isinstance(first, AffineBC) and np.array_equal(b.a_l, first.a_l) and
                                      np.array_equal(b.a_r, first.a_r) and b.b_l == first.b_l and b.b_r == first.b_r)
                if not same:
                    raise NotImplementedError("operators on one axis with different boundary conditions must be applied separately")
            bcs.append(first)
        super().__init__([(L, L.axis - 1) for L, _ in terms], bcs, tuple(global_shape), [False] * nd, dtype,
                         flags=flags, dist=ctx._h if ctx is not None else None,
                         local_rank=None if ctx is not None else local_rank)
        self.ctx = ctx
        rank, nranks = (ctx.rank, ctx.nranks) if ctx is not None else local_rank
        self.start, self.count = slab_bounds(global_shape[-1], nranks, rank)
        h = C.c_int32(0)
        _lib.check(_lib.load().deo_dist_plan_halo(self._h, C.byref(h)))
        self.halo = int(h.value)
        self.global_shape = tuple(global_shape)
        self.local_in_shape = tuple(global_shape[:-1]) + (self.count + 2 * self.halo,)
        self.local_out_shape = tuple(global_shape[:-1]) + (self.count,)

    def apply(self, du: DeviceArray, u_ext: DeviceArray):
        _lib.check(_lib.load().deo_dist_plan_apply(self._h, du._h, u_ext._h))

    def apply_host(self, du_host: np.ndarray, u_ext_host: np.ndarray):
        """Host-buffer form on a slab: `u_ext_host` is this rank's [halo | own planes | halo] block (only the own planes
        are read), `du_host` its `count` output planes."""
        own = u_ext_host[..., self.halo:self.halo + self.count] if self.halo else u_ext_host
        assert own.flags.f_contiguous and du_host.flags.f_contiguous
        _lib.check(_lib.load().deo_dist_plan_apply_host(self._h, du_host.ctypes.data_as(C.c_void_p), own.ctypes.data_as(C.c_void_p)))

    def time(self, du: DeviceArray, u_ext: DeviceArray, reps: int) -> float:
        ms = C.c_float(0)
        _lib.check(_lib.load().deo_dist_plan_time(self._h, du._h, u_ext._h, reps, C.byref(ms)))
        return float(ms.value)


# ---- host-side scatter / gather over the host runtime's process group (no arithmetic) ---------------------
def _np_to_torch(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a.reshape(-1, order="F")))


def broadcast_unique_id():
    """Rank 0 creates the NCCL unique id, every rank returns the same 128 bytes (any torch backend)."""
    import torch
    import torch.distributed as dist
    backend = dist.get_backend()
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    t = torch.zeros(_lib.DEO_DIST_ID_BYTES, dtype=torch.uint8, device=dev)
    if dist.get_rank() == 0:
        t.copy_(torch.frombuffer(bytearray(SlabContext.new_unique_id()), dtype=torch.uint8))
    dist.broadcast(t, src=0)
    return bytes(t.cpu().numpy().tobytes())


def scatter_slabs(u_global, global_shape, dtype, halo: int, fill=0.0):
    """deo_dist_scatter of the proposed ABI, done by the host runtime: rank 0 holds `u_global`; every rank
    returns its [halo | own planes | halo] host block with the halo planes already filled from the
    neighbouring slabs (physical faces get `fill`, they are never read)."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    dtype = np.dtype(dtype)
    plane = int(np.prod(global_shape[:-1], dtype=np.int64))
    s, c = slab_bounds(global_shape[-1], world, rank)
    mine = torch.empty(plane * (c + 2 * halo), dtype=torch.from_numpy(np.empty(0, dtype)).dtype)
    if rank == 0:
        for r in range(world):
            ext = extended_slab(np.asarray(u_global, dtype=dtype), r, world, halo, fill=fill)
            t = _np_to_torch(ext)
            if r == 0:
                mine.copy_(t)
            else:
                dist.send(t, dst=r)
    else:
        dist.recv(mine, src=0)
    return mine.numpy().reshape(tuple(global_shape[:-1]) + (c + 2 * halo,), order="F")


def gather_slabs(du_local, global_shape):
    """deo_dist_gather: rank 0 returns the reassembled global array, the other ranks None."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    if rank != 0:
        dist.send(_np_to_torch(du_local), dst=0)
        return None
    out = np.empty(tuple(global_shape), dtype=du_local.dtype, order="F")
    plane = int(np.prod(global_shape[:-1], dtype=np.int64))
    for r in range(world):
        s, c = slab_bounds(global_shape[-1], world, r)
        if r == 0:
            out[..., s:s + c] = du_local
        else:
            t = torch.empty(plane * c, dtype=_np_to_torch(du_local).dtype)
            dist.recv(t, src=r)
            out[..., s:s + c] = t.numpy().reshape(tuple(global_shape[:-1]) + (c,), order="F")
    return out


def exchange_halos_host(ext, halo: int):
    """The library's exchange protocol (dist.cu: send my first/last `halo` own planes, receive into my
    halo planes) replayed on host arrays over the process group -- used by the CPU tests of the N>1 path."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    cnt = ext.shape[-1] - 2 * halo
    reqs, bufs = [], []
    for nb, send_sl, recv_sl in ((rank - 1, slice(halo, 2 * halo), slice(0, halo)),
                                 (rank + 1, slice(cnt, cnt + halo), slice(cnt + halo, cnt + 2 * halo))):
        if nb < 0 or nb >= world or halo == 0:
            continue
        st = _np_to_torch(ext[..., send_sl])
        rt = torch.empty_like(st)
        reqs += [dist.isend(st, dst=nb), dist.irecv(rt, src=nb)]
        bufs.append((recv_sl, rt))
    for r in reqs:
        r.wait()
    for sl, rt in bufs:
        ext[..., sl] = rt.numpy().reshape(ext[..., sl].shape, order="F")
    return ext
