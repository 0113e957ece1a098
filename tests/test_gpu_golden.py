"""GPU vs the committed golden vectors (tests/golden/knn_golden.npz, minted by the oracle): ids identical,
distances within 1e-5 relative.  Both search paths (scan kernel, tensor-core kernel) are checked."""
import os

import numpy as np
import pytest

from oracle import cpu_ref as R
from tests.golden.gen_golden import CASES, case_inputs

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("path", [1, 2], ids=["scan", "tensor"])
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_cuda_matches_golden(vdb, case, path):
    name, metric, dim, store, n, nq, k, scale, n_del = case
    gold = np.load(os.path.join(GOLD, "knn_golden.npz"))
    raw, q, deleted = case_inputs(metric, dim, store, n, nq, scale, n_del)
    ix = vdb.Index(metric, dim, store_dtype=store)
    ix.init_index(n)
    ix.add_items(raw, np.arange(n))
    if deleted:
        ix.mark_deleted(deleted)
    ix.set_option("path", path)
    labels, dist, cnt = ix.knn_query_padded(q, k)
    want_l, want_d, want_c = gold[name + "/labels"], gold[name + "/dist"], gold[name + "/counts"]
    assert np.array_equal(cnt, want_c)
    stored = R.prepare_rows(raw, metric, store)
    for i in range(nq):
        c = int(cnt[i])
        if not np.array_equal(labels[i, :c], want_l[i, :c]):       # only ties within tolerance may differ
            msg = R.check_topk(labels[i, :c], dist[i, :c], q[i], stored, np.arange(n), k, metric, deleted=deleted, rtol=1e-5)
            assert msg is None, msg
        np.testing.assert_allclose(dist[i, :c], want_d[i, :c], rtol=1e-5, atol=1e-6)
        assert (labels[i, c:] == -1).all()
