"""Device / pinned-host arrays owned through the C ABI (fc_device_malloc & co)."""
import ctypes as C

import numpy as np

from ._lib import lib, check


class DeviceArray:
    """n doubles in HBM on `device`."""

    def __init__(self, n, device=0):
        self.n = int(n)
        self.device = int(device)
        p = C.c_void_p()
        check(lib.fc_device_malloc(self.device, self.n * 8, C.byref(p)))
        self.ptr = p.value
        self._owned = True

    @classmethod
    def from_numpy(cls, a, device=0):
        a = np.ascontiguousarray(a, dtype=np.float64)
        d = cls(a.size, device)
        d.upload(a)
        return d

    def upload(self, a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        assert a.size == self.n
        if self.n:
            check(lib.fc_memcpy_h2d(self.device, self.ptr, a.ctypes.data, self.n * 8))

    def download(self, out=None):
        if out is None:
            out = np.empty(self.n, dtype=np.float64)
        if self.n:
            check(lib.fc_memcpy_d2h(self.device, out.ctypes.data, self.ptr, self.n * 8))
        return out

    def free(self):
        if self._owned and self.ptr:
            lib.fc_device_free(self.device, self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def pinned_empty(n):
    """float64 NumPy array of n elements backed by page-locked host memory (cudaHostAlloc)."""
    n = int(n)
    p = C.c_void_p()
    check(lib.fc_host_malloc_pinned(max(n, 1) * 8, C.byref(p)))
    buf = (C.c_double * max(n, 1)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=np.float64, count=n)
    _PINNED[arr.ctypes.data] = p.value
    return arr


_PINNED = {}


def pinned_free(arr):
    p = _PINNED.pop(arr.ctypes.data, None)
    if p:
        lib.fc_host_free_pinned(p)


def addr_of(a):
    return a.ptr if isinstance(a, DeviceArray) else a.ctypes.data
