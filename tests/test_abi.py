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


def _level_plan(vdb, nq, n_rows, k):
    lib = vdb._ffi.lib()
    buf = (C.c_int * 256)()
    n = lib.vdb_debug_level_plan(nq, n_rows, k, buf, 256)
    assert 10 <= n <= 256
    v = list(buf[:n])
    keys = ("kp", "kq", "cap", "growth", "query_blocks", "n_tiles", "n_pos", "probe_tiles", "probe_rank")
    plan = dict(zip(keys, v[:9]))
    plan["levels"] = [tuple(v[10 + 3 * i: 13 + 3 * i]) for i in range(v[9])]
    return plan


def test_level_plan_invariants_over_a_grid_of_shapes(vdb):
    """host arithmetic of the batched search (csrc/gemm_topk.cu gemm_topk_level_plan): every tile position is
    visited by exactly one threshold level, ranks and buffers are consistent, shared-memory needs stay in budget."""
    for k in (1, 5, 10, 16, 17, 32, 33, 64, 100, 128):
        for n_rows in (1, 255, 256, 257, 1000, 1024, 1025, 4096, 5000, 8192, 9000, 40_000, 125_000, 1_000_000,
                       1_250_000, 12_500_000, 100_000_000, 4_000_000_000):
            for nq in (5, 256, 512, 513, 1024, 4096, 8192, 100_000):
                p = _level_plan(vdb, nq, n_rows, k)
                what = f"k={k} n_rows={n_rows} nq={nq}: {p}"
                kp, kq, cap = p["kp"], p["kq"], p["cap"]
                assert kp >= 2 * k and kp >= 32 and kp & (kp - 1) == 0 and kp <= 256, what
                assert k <= kq <= kp and kq >= kp // 2, what
                assert p["n_tiles"] == (n_rows + 255) // 256 and p["n_pos"] >= p["n_tiles"] > p["n_pos"] // 2 or p["n_pos"] == 1, what
                assert p["query_blocks"] == (nq + 255) // 256, what
                # buffers: a power of two, at least 16 k', at most 8192 keys (shared memory of select / window re-rank:
                # (8192 + k') * 8 <= 80 KB and 2 * 8192 * 8 = 128 KB <= 227 KB); shards up to 32 k' rows fit entirely
                assert cap & (cap - 1) == 0 and 16 * kp <= cap <= 8192, what
                if n_rows <= 32 * kp:
                    assert cap >= n_rows, what
                # probe: 8 chunk minima per position fit the buffer and (when the shard is large enough) are at
                # least 4x the rank
                P = p["probe_tiles"]
                assert 1 <= P <= p["n_pos"] and 8 * P <= cap, what
                assert p["probe_rank"] in (kq, kp), what
                if p["n_pos"] >= 4 * P:
                    assert p["probe_rank"] == kq and 8 * P >= 4 * kq, what
                # levels tile [0, n_pos) in order; every level but the last publishes a threshold rank
                lv = p["levels"]
                assert lv and lv[0][0] == 0 and lv[-1][1] == p["n_pos"] and lv[-1][2] == 0, what
                assert len(lv) <= 12, what
                seen, rank = P, p["probe_rank"]
                for i, (p0, p1, r_after) in enumerate(lv):
                    assert p1 > p0 and (i == 0 or p0 == lv[i - 1][1]), what
                    # expected survivors (rank x positions of the level / positions behind the threshold) + k' carried
                    # stay within 55 % of the buffer whenever the threshold exists at all (shards of <= 32 k' rows
                    # may run without one: their buffer holds every row)
                    if n_rows > 32 * kp:
                        assert rank * (p1 - p0) / seen + kp <= 0.55 * cap + 1 or (p1 - p0) <= seen * p["growth"], what
                        assert rank * min(p1 - p0, seen * p["growth"]) / seen + kp <= cap, what
                    if i < len(lv) - 1:
                        assert r_after == (kq if 4 * p1 <= p["n_pos"] else kp), what
                        seen, rank = p1, r_after
                # small batches: probe + at most 2 levels up to 16M rows
                if p["query_blocks"] <= 2 and k <= 16 and n_rows <= 16_000_000:
                    assert len(lv) <= 2, what
    lib = vdb._ffi.lib()
    assert lib.vdb_debug_level_plan(0, 10, 10, None, 0) == vdb._ffi.VDB_EINVAL
    assert lib.vdb_debug_level_plan(10, 10, 129, None, 0) == vdb._ffi.VDB_EINVAL
    assert lib.vdb_debug_level_plan(10, 1_000_000, 10, None, 0) >= 13          # size query
