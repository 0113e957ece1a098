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
g = np.sort(np.random.default_rng(0).random((4, 9, 10), dtype=np.float32), axis=2)
ids = np.arange(4 * 9 * 10, dtype=np.int64).reshape(4, 9, 10)
vdb.merge_topk(g, ids, 10)
print("sanity_small ok,", vdb.launch_count(), "launches")
