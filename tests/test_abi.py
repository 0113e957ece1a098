"""CPU: the C-ABI shared library loads, exports every symbol include/vdb.h declares, and -- with no GPU in
this container -- every compute entry point fails loudly instead of falling back to a CPU path."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "vdb.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vdb_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(vdb):
    lib = vdb._ffi.lib()
    names = header_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"libvdb_b200.so does not export {name}"
    assert set(names) == set(vdb._ffi.SIGNATURES), "ctypes prototypes out of sync with include/vdb.h"
    assert b"sm_100a" in lib.vdb_version()


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback(vdb):
    ix = vdb.Index("cosine", 512)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ix.init_index(100)
    with pytest.raises(RuntimeError):
        ix.knn_query(np.zeros((1, 512), np.float32), 1)          # not initialised
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        vdb.merge_topk(np.zeros((2, 1, 4), np.float32), np.zeros((2, 1, 4), np.int64), 4)


def test_argument_validation_without_gpu(vdb):
    lib = vdb._ffi.lib()
    h = C.c_void_p()
    assert lib.vdb_create(0, 0, 0, 10, 0, C.byref(h)) == vdb._ffi.VDB_EINVAL
    assert lib.vdb_create(512, 7, 0, 10, 0, C.byref(h)) == vdb._ffi.VDB_EINVAL
    assert lib.vdb_create(512, 0, 5, 10, 0, C.byref(h)) == vdb._ffi.VDB_EINVAL
    assert b"metric" in lib.vdb_last_error() or b"store_dtype" in lib.vdb_last_error()
    assert lib.vdb_count(None) == 0 and lib.vdb_capacity(None) == 0
    with pytest.raises(RuntimeError, match="Space name"):
        vdb.Index("hamming", 512)
    assert lib.vdb_launch_count() == 0 or _has_gpu()


def test_oracle_is_not_imported_by_the_product():
    """only tests/, smoke() and bench.py's CPU legs may touch oracle/."""
    pkg = os.path.join(ROOT, "distributed-vector-database_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), encoding="utf-8").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f"{f} imports oracle"
                assert "knn_ref" not in text, f"{f} references the oracle's C port"
