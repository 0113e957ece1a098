"""ctypes loader for oracle/hnsw_ref.c -- TEST / BENCH INFRASTRUCTURE ONLY.

A restatement of hnswlib v0.8.0's HNSW index (the reference's search structure: `hnswlib.Index`, M=32,
ef_construction=128, ef=max(50, 2k) -- src/datanode/handler.py:86,360-364).  hnswlib itself is third-party, not vendored
by the reference and not installable here.  This port answers one question for `bench.py --impl reference`: how many
queries per second, at what recall, the reference's APPROXIMATE walk does on the box's cores.  It is not a parity
oracle (the oracle is the exact scan in cpu_ref.py / knn_ref.c)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libhnsw_ref.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "hnsw_ref.c")
        if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
            subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "_build/libhnsw_ref.so"])
        L = C.CDLL(_SO)
        fp = C.POINTER(C.c_float)
        L.hnsw_build.argtypes = [fp, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_int]
        L.hnsw_build.restype = C.c_void_p
        L.hnsw_search.argtypes = [C.c_void_p, fp, C.c_size_t, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int64), fp]
        L.hnsw_search.restype = None
        L.hnsw_max_level.argtypes = [C.c_void_p]
        L.hnsw_mean_degree0.argtypes = [C.c_void_p]
        L.hnsw_mean_degree0.restype = C.c_double
        L.hnsw_free.argtypes = [C.c_void_p]
        L.hnsw_free.restype = None
        _lib = L
    return _lib


class HnswPort:
    """`stored`: float32 [n, dim] rows AS THE INDEX HOLDS THEM (normalised for cosine), kept alive by this object.
    metric "l2" -> squared L2; "ip" / "cosine" -> 1 - dot."""

    def __init__(self, stored: np.ndarray, metric: str, M: int = 32, ef_construction: int = 128, seed: int = 100,
                 nthreads: int = 0):
        self.stored = np.ascontiguousarray(stored, dtype=np.float32)
        self.metric = metric
        n, dim = self.stored.shape
        self._h = lib().hnsw_build(self.stored.ctypes.data_as(C.POINTER(C.c_float)), n, dim, 0 if metric == "l2" else 1,
                                   M, ef_construction, seed, nthreads)
        if not self._h:
            raise RuntimeError("hnsw_build failed (M in [2, 64], n > 0)")

    def knn_query(self, queries: np.ndarray, k: int, ef: int, nthreads: int = 0):
        """queries as the index would see them (normalise them yourself for cosine) -> (labels int64 [nq, k], dist)"""
        q = np.ascontiguousarray(np.atleast_2d(queries), dtype=np.float32)
        nq = q.shape[0]
        labels = np.empty((nq, k), dtype=np.int64)
        dist = np.empty((nq, k), dtype=np.float32)
        lib().hnsw_search(self._h, q.ctypes.data_as(C.POINTER(C.c_float)), nq, k, ef, nthreads,
                          labels.ctypes.data_as(C.POINTER(C.c_int64)), dist.ctypes.data_as(C.POINTER(C.c_float)))
        return labels, dist

    @property
    def max_level(self) -> int:
        return int(lib().hnsw_max_level(self._h))

    @property
    def mean_degree0(self) -> float:
        return float(lib().hnsw_mean_degree0(self._h))

    def close(self):
        if self._h:
            lib().hnsw_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
