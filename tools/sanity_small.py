"""Small end-to-end exercise of every kernel (scan, tensor path with both operand planes, tombstones,
insert, merge) for compute-sanitizer runs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import dvdb_b200 as vdb
from oracle import cpu_ref as R

for metric, store, dim in (("cosine", "f32", 512), ("l2", "f32", 100), ("ip", "f16", 768)):
    n = 3000
    raw = R.synth_rows(R.SEED_DB, 0, n, dim)
    ix = vdb.Index(metric, dim, store_dtype=store)
    ix.init_index(n + 10)
    ix.add_items(raw, np.arange(n))
    ix.mark_deleted([5, 77])
    stored = R.prepare_rows(raw, metric, store)
    q = R.synth_rows(R.SEED_QUERY, 0, 40, dim)
    for path, shadow in ((1, 1), (2, 1), (2, 0)):
        ix.set_option("path", path)
        ix.set_option("shadow", shadow)
        l, d, c = ix.knn_query_padded(q if path == 2 else q[:3], 10)
        for i in range(len(l)):
            msg = R.check_topk(l[i], d[i], q[i], stored, np.arange(n), 10, metric, deleted=[5, 77], rtol=1e-5)
            assert msg is None, (metric, store, path, shadow, i, msg)
    ix.close()
# round 2 paths: single-query scans whose last CTAs merge the lists (one level at k = 10, two levels at k = 100 over enough
# rows for a full grid), the device-side exact fallback (near-duplicate flood: certificates fail), growth in place past the
# address reservation while rows are present, non-affine labels
n, dim = 40_000, 512
ix = vdb.Index("l2", dim)
ix.init_index(n)
ix.add_synthetic(R.SEED_DB, 0, n)
stored = R.synth_rows(R.SEED_DB, 0, n, dim)
q = R.synth_rows(R.SEED_QUERY, 0, 2, dim)
for k, nq in ((10, 1), (100, 1), (100, 2), (300, 1)):
    l, d, c = ix.knn_query_padded(q[:nq], k)
    for i in range(nq):
        msg = R.check_topk(l[i], d[i], q[i], stored, np.arange(n), k, "l2", rtol=1e-5)
        assert msg is None, ("fused merge", k, nq, i, msg)
ix.resize_index(9_000_000)                                   # beyond the reservation: re-base, content intact
l2, d2, _ = ix.knn_query_padded(q[:1], 10)
assert np.array_equal(l2[0], ix.knn_query_padded(q[:1], 10)[0][0])
ix.close()
base = R.synth_rows(R.SEED_DB, 0, 1, dim)[0]
dup = np.tile(base, (3000, 1)) + np.random.default_rng(1).normal(0, 1e-6, (3000, dim)).astype(np.float32)
ix = vdb.Index("cosine", dim)
ix.init_index(3000)
ix.add_items(dup, np.arange(3000)[::-1].copy())             # labels in reverse: not base + row
ix.set_option("path", 2)
qd = R.synth_rows(R.SEED_QUERY, 0, 12, dim)
l, d, c = ix.knn_query_padded(qd, 10)
st = R.prepare_rows(dup, "cosine")
for i in range(len(qd)):
    msg = R.check_topk(l[i], d[i], qd[i], st, np.arange(3000)[::-1], 10, "cosine", rtol=1e-5)
    assert msg is None, ("fallback", i, msg)
assert ix.get_stat("fallback_queries") > 0, "the near-duplicate flood was expected to fail the certificates"
ix.close()
g = np.sort(np.random.default_rng(0).random((4, 9, 10), dtype=np.float32), axis=2)
ids = np.arange(4 * 9 * 10, dtype=np.int64).reshape(4, 9, 10)
vdb.merge_topk(g, ids, 10)
print("sanity_small ok,", vdb.launch_count(), "launches")
