"""ctypes loader for oracle/knn_ref.c -- TEST INFRASTRUCTURE ONLY (see cpu_ref.py header)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libknn_ref.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "knn_ref.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "_build/libknn_ref.so"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        fp, i64p, u8p, ip = (C.POINTER(C.c_float), C.POINTER(C.c_int64), C.POINTER(C.c_uint8),
                             C.POINTER(C.c_int))
        L.oracle_knn.argtypes = [fp, C.c_size_t, fp, i64p, C.c_size_t, C.c_int, C.c_size_t, C.c_int,
                                 u8p, C.c_int, i64p, fp, ip, C.c_int]
        L.oracle_knn.restype = C.c_int
        L.oracle_distances.argtypes = [fp, fp, C.c_size_t, C.c_int, C.c_size_t, C.c_int, fp]
        L.oracle_normalize.argtypes = [fp, C.c_size_t, C.c_int, fp]
        L.oracle_synth_rows.argtypes = [C.c_uint64, C.c_uint64, C.c_size_t, C.c_int, fp]
        L.oracle_num_threads.restype = C.c_int
        L.oracle_set_threads.argtypes = [C.c_int]
        L.oracle_round_f16.argtypes = [fp, C.c_size_t]
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def synth_rows(seed: int, row_start: int, n: int, dim: int) -> np.ndarray:
    out = np.empty((n, dim), dtype=np.float32)
    lib().oracle_synth_rows(seed, row_start, n, dim, _p(out, C.c_float))
    return out


def normalize(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(np.atleast_2d(x), dtype=np.float32)
    out = np.empty_like(x)
    lib().oracle_normalize(_p(x, C.c_float), x.shape[0], x.shape[1], _p(out, C.c_float))
    return out


def distances(q: np.ndarray, stored: np.ndarray, metric: str) -> np.ndarray:
    q = np.ascontiguousarray(q, dtype=np.float32).reshape(-1)
    if metric == "cosine":
        q = normalize(q)[0]
    stored = np.ascontiguousarray(stored, dtype=np.float32)
    out = np.empty(stored.shape[0], dtype=np.float32)
    lib().oracle_distances(_p(q, C.c_float), _p(stored, C.c_float), stored.shape[0], stored.shape[1],
                           stored.shape[1], 0 if metric == "l2" else 1, _p(out, C.c_float))
    return out


def knn(queries: np.ndarray, stored: np.ndarray, labels, k: int, metric: str, dead=None,
        nthreads: int = 0):
    """Exact top-k (ascending (distance,label)) of ``queries`` over prepared rows ``stored``."""
    queries = np.ascontiguousarray(np.atleast_2d(queries), dtype=np.float32)
    if metric == "cosine":
        queries = normalize(queries)
    stored = np.ascontiguousarray(stored, dtype=np.float32)
    labels = None if labels is None else np.ascontiguousarray(labels, dtype=np.int64)
    dead = None if dead is None else np.ascontiguousarray(dead, dtype=np.uint8)
    nq = queries.shape[0]
    out_l = np.empty((nq, k), dtype=np.int64)
    out_d = np.empty((nq, k), dtype=np.float32)
    cnt = np.empty(nq, dtype=np.int32)
    rc = lib().oracle_knn(_p(queries, C.c_float), nq, _p(stored, C.c_float), _p(labels, C.c_int64),
                          stored.shape[0], stored.shape[1], stored.shape[1],
                          0 if metric == "l2" else 1, _p(dead, C.c_uint8), k,
                          _p(out_l, C.c_int64), _p(out_d, C.c_float), _p(cnt, C.c_int), nthreads)
    if rc != 0:
        raise RuntimeError(f"oracle_knn failed rc={rc}")
    return out_l, out_d, cnt


def set_threads(n: int) -> None:
    """Thread count of every OpenMP region of the C oracle (overrides an inherited OMP_NUM_THREADS)."""
    lib().oracle_set_threads(int(n))


def host_cores() -> int:
    """Cores this process may run on (the CPU baseline states this number)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def round_f16_(x: np.ndarray) -> np.ndarray:
    """In place fp32 -> fp16 -> fp32 (what an fp16-stored shard holds)."""
    assert x.dtype == np.float32 and x.flags.c_contiguous
    lib().oracle_round_f16(_p(x, C.c_float), x.size)
    return x


def num_threads() -> int:
    return int(lib().oracle_num_threads())
