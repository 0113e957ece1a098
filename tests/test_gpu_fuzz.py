"""GPU differential test: the batched tensor-core path (probe, levels, selects, window re-rank) against the
single-query scan path -- different kernels, same summation order in the exact distances -- over random shapes.
ids and distances must agree bit for bit; the scan path itself is pinned to the CPU oracle in test_gpu_search.py."""
import numpy as np
import pytest

from oracle import cpu_ref as R

pytestmark = pytest.mark.gpu


def _case(rng):
    dim = int(rng.choice([32, 64, 100, 128, 256, 512, 768]))
    store = "f16" if rng.random() < 0.3 else "f32"
    metric = str(rng.choice(["l2", "ip", "cosine"]))
    n = int(rng.choice([1, 37, 300, 513, 1025, 2049, 4097, 5000, 9000, 17000, 33000, 70000]))
    k = int(rng.choice([1, 3, 10, 16, 17, 40, 64, 100, 128]))
    nq = int(rng.choice([5, 9, 64, 200, 257, 600, 1100]))
    dead_frac = float(rng.choice([0.0, 0.0, 0.05, 0.5]))
    shadow = int(rng.random() < 0.8)
    return dim, store, metric, n, k, nq, dead_frac, shadow


@pytest.mark.parametrize("seed", range(96))
def test_tensor_path_equals_scan_path_on_random_shapes(vdb, seed):
    rng = np.random.default_rng(1000 + seed)
    dim, store, metric, n, k, nq, dead_frac, shadow = _case(rng)
    ix = vdb.Index(metric, dim, store_dtype=store)
    ix.init_index(n)
    ix.add_synthetic(R.SEED_DB + seed, 0, n)
    if dead_frac and n > 1:
        dead = np.flatnonzero(rng.random(n) < dead_frac)
        if len(dead):
            ix.mark_deleted(dead.tolist())
    if store == "f32":
        ix.set_option("shadow", shadow)
    q = R.synth_rows(R.SEED_QUERY + seed, 0, nq, dim)
    if seed % 3 == 0 and n > 8:                      # some queries ARE rows (distance 0, exact ties with nothing)
        q[:4] = R.synth_rows(R.SEED_DB + seed, 0, 4, dim)
    ix.set_option("path", 2)
    lt, dt, ct = ix.knn_query_padded(q, k)
    tb = ix.get_stat("tensor_batches")
    ix.set_option("path", 1)
    m = min(nq, 48)
    ls, ds, cs = ix.knn_query_padded(q[:m], k)
    what = f"dim={dim} store={store} metric={metric} n={n} k={k} nq={nq} dead={dead_frac} shadow={shadow}"
    assert tb >= 1, what
    assert np.array_equal(ct[:m], cs), what
    assert np.array_equal(lt[:m], ls), what
    assert np.array_equal(dt[:m].view(np.uint32), ds.view(np.uint32)), what
    assert (np.diff(dt[:, :max(int(ct.min()), 1)], axis=1) >= 0).all(), what
