"""TEST INFRASTRUCTURE ONLY -- scipy.spatial.cKDTree restated with its exact tie behaviour
(``ckdtree_oracle.c``): ``KDTree(data).query(x, k)`` returns the same neighbours IN THE SAME ORDER
as ``cKDTree(data).query(x, k)`` for 2-D data, equidistant and coincident points included."""
import ctypes

import numpy as np

from . import lib

_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int64)


class KDTree:
    def __init__(self, data):
        self.data = np.ascontiguousarray(data, dtype=np.float64)
        if self.data.ndim != 2 or self.data.shape[1] != 2:
            raise ValueError("2-D points (n, 2) only")
        self.n = self.data.shape[0]
        L = lib()
        L.ora_kd_build.restype = ctypes.c_void_p
        L.ora_kd_build.argtypes = [_dp, ctypes.c_int64]
        L.ora_kd_free.argtypes = [ctypes.c_void_p]
        L.ora_kd_error.argtypes = [ctypes.c_void_p]
        L.ora_kd_indices.argtypes = [ctypes.c_void_p, _ip]
        L.ora_kd_query_many.argtypes = [ctypes.c_void_p, _dp, ctypes.c_int64, ctypes.c_int64, _ip, _dp]
        self._L = L
        self._t = L.ora_kd_build(self.data.ctypes.data_as(_dp), self.n)
        if L.ora_kd_error(self._t):
            raise NotImplementedError("introselect depth limit reached (heap_select path not restated)")

    def __del__(self):
        if getattr(self, "_t", None):
            self._L.ora_kd_free(self._t)
            self._t = None

    @property
    def indices(self):
        out = np.empty(self.n, dtype=np.int64)
        self._L.ora_kd_indices(self._t, out.ctypes.data_as(_ip))
        return out

    def query(self, x, k=1):
        xs = np.ascontiguousarray(np.atleast_2d(np.asarray(x, dtype=np.float64)))
        nq = xs.shape[0]
        idx = np.empty((nq, k), dtype=np.int64)
        dist = np.empty((nq, k), dtype=np.float64)
        self._L.ora_kd_query_many(self._t, xs.ctypes.data_as(_dp), nq, k, idx.ctypes.data_as(_ip),
                                  dist.ctypes.data_as(_dp))
        if k == 1:
            dist, idx = dist[:, 0], idx[:, 0]
        if np.ndim(x) == 1:
            return dist[0], idx[0]
        return dist, idx
