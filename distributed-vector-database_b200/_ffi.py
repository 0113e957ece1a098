"""ctypes binding of the C ABI in include/vdb.h (libvdb_b200.so, built in-tree by csrc/build.sh).

There is no fallback: if the shared library is missing the import of any compute entry point
raises, and on a box without a CUDA device every call returns VDB_ECUDA -> RuntimeError."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvdb_b200.so")

VDB_L2, VDB_IP, VDB_COSINE = 0, 1, 2
VDB_F32, VDB_F16 = 0, 1
VDB_OK, VDB_EINVAL, VDB_ECUDA, VDB_EFULL, VDB_ENOMEM, VDB_EIO, VDB_ENOTFOUND = 0, -1, -2, -3, -4, -5, -6
METRIC_CODES = {"l2": VDB_L2, "ip": VDB_IP, "cosine": VDB_COSINE}
DTYPE_CODES = {"f32": VDB_F32, "float32": VDB_F32, "f16": VDB_F16, "float16": VDB_F16}

_f32p = C.POINTER(C.c_float)
_i64p = C.POINTER(C.c_int64)
_i32p = C.POINTER(C.c_int)
_vp = C.c_void_p

# every symbol include/vdb.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "vdb_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_int, C.POINTER(_vp)]),
    "vdb_destroy": (None, [_vp]),
    "vdb_add": (C.c_int, [_vp, _f32p, _i64p, C.c_size_t]),
    "vdb_add_dev": (C.c_int, [_vp, _vp, _i64p, C.c_size_t, _vp]),
    "vdb_add_synthetic": (C.c_int, [_vp, C.c_uint64, C.c_uint64, C.c_size_t, C.c_int64]),
    "vdb_synth_dev": (C.c_int, [C.c_uint64, C.c_uint64, C.c_size_t, C.c_int, _vp, _vp]),
    "vdb_mark_deleted": (C.c_int, [_vp, _i64p, C.c_size_t]),
    "vdb_unmark_deleted": (C.c_int, [_vp, _i64p, C.c_size_t]),
    "vdb_search": (C.c_int, [_vp, _f32p, C.c_size_t, C.c_int, _i64p, _f32p, _i32p]),
    "vdb_search_submit": (C.c_int, [_vp, _f32p, C.c_size_t, C.c_int, _i64p, _f32p, _i32p, C.POINTER(_vp)]),
    "vdb_search_collect": (C.c_int, [_vp]),
    "vdb_host_alloc": (_vp, [C.c_size_t]),
    "vdb_host_free": (None, [_vp]),
    "vdb_search_dev": (C.c_int, [_vp, _vp, C.c_size_t, C.c_int, _vp, _vp, _vp, _vp]),
    "vdb_count": (C.c_size_t, [_vp]),
    "vdb_live_count": (C.c_size_t, [_vp]),
    "vdb_capacity": (C.c_size_t, [_vp]),
    "vdb_dim": (C.c_int, [_vp]),
    "vdb_resize": (C.c_int, [_vp, C.c_size_t]),
    "vdb_get_rows": (C.c_int, [_vp, _i64p, C.c_size_t, _f32p]),
    "vdb_save": (C.c_int, [_vp, C.c_char_p]),
    "vdb_load": (C.c_int, [C.c_char_p, C.c_size_t, C.c_int, C.POINTER(_vp)]),
    "vdb_save_image": (C.c_int, [_vp, C.c_char_p, C.c_char_p]),
    "vdb_load_image": (C.c_int, [C.c_char_p, C.c_char_p, C.c_size_t, C.c_int, C.POINTER(_vp)]),
    "vdb_merge_topk": (C.c_int, [_vp, _vp, C.c_int, C.c_size_t, C.c_int, C.c_int, _vp, _vp, C.c_int, C.c_int, _vp]),
    "vdb_xchg_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_int, C.POINTER(_vp), _vp]),
    "vdb_xchg_connect": (C.c_int, [_vp, _vp]),
    "vdb_xchg_create_q": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_size_t, C.POINTER(_vp), _vp]),
    "vdb_xchg_query_slot": (_vp, [_vp, C.c_int]),
    "vdb_xchg_gather_queries": (C.c_int, [_vp, _vp, C.c_size_t, C.c_int, C.c_uint32, _vp]),
    "vdb_xchg_wait_queries": (C.c_int, [_vp, C.c_int, C.c_uint32, _vp]),
    "vdb_xchg_merge_dev": (C.c_int, [_vp, _vp, _vp, C.c_size_t, C.c_int, _vp, _vp, _vp]),
    "vdb_xchg_status": (C.c_int, [_vp]),
    "vdb_xchg_destroy": (None, [_vp]),
    "vdb_launch_count": (C.c_uint64, []),
    "vdb_set_option": (C.c_int, [_vp, C.c_char_p, C.c_long]),
    "vdb_get_stat": (C.c_long, [_vp, C.c_char_p]),
    "vdb_last_error": (C.c_char_p, []),
    "vdb_version": (C.c_char_p, []),
    "vdb_debug_level_plan": (C.c_int, [C.c_size_t, C.c_size_t, C.c_int, C.POINTER(C.c_int), C.c_int]),
}

_lib = None


def lib():
    """Load libvdb_b200.so (once) and attach the prototypes.  Raises if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)          # AttributeError if the .so does not export it
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def last_error() -> str:
    msg = lib().vdb_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc: int, what: str = "") -> None:
    """hnswlib raises RuntimeError for every index-level failure (the reference catches exactly
    that at src/datanode/handler.py:272,366); so does this binding."""
    if rc != VDB_OK:
        raise RuntimeError(f"{what + ': ' if what else ''}{last_error()} (vdb error {rc})")
