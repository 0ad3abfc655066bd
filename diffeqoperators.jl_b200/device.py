"""DeviceArray: the handle type the Julia glue exposes (`DeviceArray{T,N} <: AbstractArray{T,N}`,
an opaque library-owned device buffer plus dims).  Column-major like Julia arrays."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


class DeviceArray:
    def __init__(self, shape, dtype, handle=None):
        self.shape = tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)
        _lib.dtype_code(self.dtype)
        self.ndim = len(self.shape)
        self.size = int(np.prod(self.shape, dtype=np.int64))
        self.nbytes = self.size * self.dtype.itemsize
        if handle is None:
            h = C.c_void_p()
            _lib.check(_lib.load().deo_buffer_create(self.nbytes, C.byref(h)))
            handle = h
        self._h = handle

    def __del__(self):
        try:
            if self._h is not None:
                _lib.load().deo_buffer_free(self._h)
                self._h = None
        except Exception:
            pass

    @classmethod
    def from_host(cls, a):
        """copyto!(DeviceArray, Array): uploads in column-major (Fortran) order."""
        a = np.asarray(a)
        d = cls(a.shape, a.dtype)
        d.upload(a)
        return d

    @classmethod
    def wrap(cls, devptr: int, shape, dtype):
        """Non-owning view of foreign device memory (e.g. torch.Tensor.data_ptr())."""
        dtype = np.dtype(dtype)
        nbytes = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
        h = C.c_void_p()
        _lib.check(_lib.load().deo_buffer_wrap(C.c_void_p(devptr), nbytes, C.byref(h)))
        return cls(shape, dtype, handle=h)

    def view(self, offset_elems: int, shape):
        """Non-owning window of this buffer (column-major): e.g. one component `A[.., n]` of a space-tensor.  The parent
        must outlive the view."""
        shape = tuple(int(v) for v in shape)
        n = int(np.prod(shape, dtype=np.int64))
        assert 0 <= offset_elems and offset_elems + n <= self.size, "view outside the buffer"
        v = DeviceArray.wrap(self.devptr + int(offset_elems) * self.dtype.itemsize, shape, self.dtype)
        v._parent = self
        return v

    def upload(self, a):
        a = np.asfortranarray(a, dtype=self.dtype)
        assert a.shape == self.shape, f"shape mismatch: {a.shape} vs {self.shape}"
        _lib.check(_lib.load().deo_buffer_upload(self._h, a.ctypes.data_as(C.c_void_p), self.nbytes))
        return self

    def to_host(self):
        """Array(::DeviceArray); synchronises."""
        out = np.empty(self.shape, dtype=self.dtype, order="F")
        _lib.check(_lib.load().deo_buffer_download(out.ctypes.data_as(C.c_void_p), self._h, self.nbytes))
        return out

    @property
    def devptr(self) -> int:
        p = C.c_void_p()
        _lib.check(_lib.load().deo_buffer_devptr(self._h, C.byref(p)))
        return int(p.value)

    def __array__(self, dtype=None, copy=None):
        a = self.to_host()
        return a if dtype is None else a.astype(dtype)


def zeros(shape, dtype=np.float64):
    return DeviceArray.from_host(np.zeros(shape, dtype=dtype, order="F"))


def sync():
    _lib.check(_lib.load().deo_sync())
