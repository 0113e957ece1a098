"""GPU parity: the CUDA search path (through the C ABI) against the CPU oracle on the same seeded
inputs.  ids must be identical to the oracle's exact result (ties by id; rows whose float64
distances tie within 1e-5 relative may swap), distances within 1e-5 relative of the definition.
Tolerance lives in oracle.cpu_ref.check_topk(rtol=1e-5)."""
import numpy as np
import pytest

from oracle import cpu_ref as R

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def build(vdb, metric, n, dim=512, store="f32", scale=1.0, seed=R.SEED_DB, cap=None):
    ix = vdb.Index(metric, dim, store_dtype=store)
    ix.init_index(cap or max(n, 1))
    raw = R.synth_rows(seed, 0, n, dim) * np.float32(scale) if n else np.zeros((0, dim), np.float32)
    if n:
        ix.add_items(raw, np.arange(n))
    return ix, raw


def assert_parity(ix, raw, metric, store, queries, k, deleted=(), labels=None):
    stored = R.prepare_rows(raw, metric, store)
    labels = np.arange(len(raw)) if labels is None else labels
    got_l, got_d, cnt = ix.knn_query_padded(queries, k)
    n_live = len(raw) - len(set(deleted))
    for i in range(len(queries)):
        assert cnt[i] == min(k, n_live)
        msg = R.check_topk(got_l[i, :cnt[i]], got_d[i, :cnt[i]], queries[i], stored, labels, k, metric,
                           deleted=deleted, rtol=RTOL)
        assert msg is None, f"query {i}: {msg}"
        assert (got_l[i, cnt[i]:] == -1).all() and np.isinf(got_d[i, cnt[i]:]).all()


@pytest.mark.parametrize("metric", ["l2", "ip", "cosine"])
@pytest.mark.parametrize("nq", [1, 2, 3, 8, 19])
def test_scan_parity_f32(vdb, metric, nq):
    ix, raw = build(vdb, metric, 5000, scale=1.7)
    q = R.synth_rows(R.SEED_QUERY, 0, nq, 512) * np.float32(0.9)
    assert_parity(ix, raw, metric, "f32", q, 10)


@pytest.mark.parametrize("metric", ["l2", "ip", "cosine"])
def test_scan_parity_f16_d768(vdb, metric):
    ix, raw = build(vdb, metric, 3000, dim=768, store="f16")
    q = R.synth_rows(R.SEED_QUERY, 0, 5, 768)
    assert_parity(ix, raw, metric, "f16", q, 10)


@pytest.mark.parametrize("k", [1, 5, 100, 257])
def test_scan_k_sizes(vdb, k):
    ix, raw = build(vdb, "l2", 4000)
    q = R.synth_rows(R.SEED_QUERY, 7, 3, 512)
    assert_parity(ix, raw, "l2", "f32", q, k)


@pytest.mark.parametrize("dim", [4, 100, 128, 130, 960])
def test_odd_dims(vdb, dim):
    ix, raw = build(vdb, "cosine", 777, dim=dim)
    q = R.synth_rows(R.SEED_QUERY, 0, 2, dim)
    assert_parity(ix, raw, "cosine", "f32", q, 7)


@pytest.mark.parametrize("n", [0, 1, 7, 15, 16, 17, 31, 33])
def test_tiny_and_ragged_shards(vdb, n):
    """empty index -> empty result (handler.py:353-354); fewer rows than k -> short result."""
    ix, raw = build(vdb, "l2", n, cap=64)
    q = R.synth_rows(R.SEED_QUERY, 0, 2, 512)
    assert_parity(ix, raw, "l2", "f32", q, 10)


def test_tombstones_and_overwrite(vdb):
    """delete = tombstone (handler.py:332); overwrite = tombstone + append (handler.py:254-264)."""
    ix, raw = build(vdb, "ip", 2000, cap=2100)
    q = R.synth_rows(R.SEED_QUERY, 0, 4, 512)
    stored = R.prepare_rows(raw, "ip")
    first, _, _ = R.knn_exact(q, stored, np.arange(2000), 10, "ip")
    dead = sorted(set(first[:, :6].reshape(-1).tolist()))          # kill most of the winners
    ix.mark_deleted(dead)
    assert ix.get_live_count() == 2000 - len(dead)
    assert ix.get_current_count() == 2000
    assert_parity(ix, raw, "ip", "f32", q, 10, deleted=dead)
    ix.unmark_deleted(dead[:3])
    assert_parity(ix, raw, "ip", "f32", q, 10, deleted=dead[3:])
    # re-adding a live label replaces its vector
    new_vec = q[0:1] * np.float32(1.0)
    ix.add_items(new_vec, [5])
    assert ix.get_current_count() == 2001
    got_l, got_d = ix.knn_query(q[0:1], 1)
    assert int(got_l[0, 0]) == 5
    np.testing.assert_allclose(ix.get_items([5])[0], new_vec[0], rtol=0, atol=0)


def test_exact_ties_break_by_label(vdb):
    """duplicate vectors: equal distances, smaller label first (hnswlib pair ordering)."""
    ix = vdb.Index("l2", 512)
    ix.init_index(64)
    base = R.synth_rows(R.SEED_DB, 0, 4, 512)
    rows = np.concatenate([base, base, base[::-1]])
    labels = np.array([40, 30, 20, 10, 41, 31, 21, 11, 12, 22, 32, 42])
    ix.add_items(rows, labels)
    got_l, got_d = ix.knn_query(base[0:1], 6)
    assert got_l[0, :3].tolist() == [40, 41, 42] and (got_d[0, :3] == 0).all()
    stored = R.prepare_rows(rows, "l2")
    want_l, want_d, _ = R.knn_exact(base[0:1], stored, labels, 6, "l2")
    assert got_l[0].astype(np.int64).tolist() == want_l[0].tolist()


def test_errors_match_hnswlib(vdb):
    ix, raw = build(vdb, "l2", 10, cap=10)
    with pytest.raises(RuntimeError):                      # full (handler.py:272 catches this)
        ix.add_items(raw[:1], [99])
    with pytest.raises(RuntimeError):                      # k > count (handler.py:366 catches this)
        ix.knn_query(raw[:1], 11)
    with pytest.raises(RuntimeError):
        ix.knn_query(np.zeros((1, 100), np.float32), 1)    # wrong dim
    with pytest.raises(RuntimeError):
        ix.mark_deleted([12345])
    ix.resize_index(20)
    ix.add_items(raw[:1], [99])
    assert ix.get_current_count() == 11 and ix.get_max_elements() == 20


def test_save_load_roundtrip(vdb, tmp_path):
    ix, raw = build(vdb, "cosine", 1500, store="f16", dim=768)
    ix.mark_deleted([3, 4, 5])
    q = R.synth_rows(R.SEED_QUERY, 0, 3, 768)
    a = ix.knn_query_padded(q, 10)
    path = str(tmp_path / "index.bin")
    ix.save_index(path)
    ix2 = vdb.Index("cosine", 768, store_dtype="f16")
    ix2.load_index(path, max_elements=4000)
    assert ix2.get_current_count() == 1500 and ix2.get_live_count() == 1497 and ix2.get_max_elements() == 4000
    b = ix2.knn_query_padded(q, 10)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def test_pinned_buffers_take_the_dma_path(vdb):
    """page-locked query / result buffers (vdb_host_alloc) give the same answer as pageable ones."""
    ix, raw = build(vdb, "cosine", 3000)
    q = R.synth_rows(R.SEED_QUERY, 0, 40, 512)
    a = ix.knn_query_padded(q, 10)
    qp = vdb.pinned_empty((40, 512), np.float32)
    qp[:] = q
    outs = (vdb.pinned_empty((40, 10), np.int64), vdb.pinned_empty((40, 10), np.float32), vdb.pinned_empty((40,), np.int32))
    b = ix.knn_query_padded(qp, 10, out=outs)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    assert b[0] is outs[0]
    with pytest.raises(RuntimeError):
        ix.knn_query_padded(qp, 10, out=(outs[0][:, :5], outs[1], outs[2]))


def test_device_synthetic_rows_bit_identical(vdb):
    """the on-device generator == the oracle's generator, so 10M-row sets can be spot-checked."""
    ix = vdb.Index("ip", 512)
    ix.init_index(70000)
    ix.add_synthetic(R.SEED_DB, 1000, 70000)
    ids = [1000, 1001, 35000, 70999]
    want = R.synth_rows(R.SEED_DB, 1000, 70000, 512)[[0, 1, 34000, 69999]]
    assert np.array_equal(ix.get_items(ids), want)


def test_merge_topk_matches_oracle(vdb):
    rng = np.random.default_rng(3)
    G, nq, k = 8, 33, 100
    dist = np.sort(rng.random((G, nq, k), dtype=np.float32), axis=2)
    dist[rng.random(dist.shape) < 0.05] = 0.25                   # exact cross-shard ties
    dist = np.sort(dist, axis=2)
    ids = rng.permutation(G * nq * k).astype(np.int64).reshape(G, nq, k)
    ids[:, :, 90:][rng.random((G, nq, 10)) < 0.3] = -1           # ragged shards
    got_d, got_i = vdb.merge_topk(dist, ids, k)
    want_d, want_i = R.merge_topk_by_id(dist, ids, k)
    assert np.array_equal(got_i, want_i) and np.array_equal(got_d, want_d)


def test_full_size_properties_1m(vdb):
    """BASELINE config 2 size (1M x 512 cosine k=10): size-independent checks -- a stored row
    queried against the index finds itself at distance ~0; results are sorted; batched and
    single-query searches agree; sampled queries equal the oracle on regenerated rows."""
    n = 1_000_000
    ix = vdb.Index("cosine", 512)
    ix.init_index(n)
    ix.add_synthetic(R.SEED_DB, 0, n)
    probe_rows = [0, 123456, 999999]
    q = R.synth_rows(R.SEED_DB, 0, 1, 512)
    q = np.concatenate([R.synth_rows(R.SEED_DB, r, 1, 512) for r in probe_rows] + [R.synth_rows(R.SEED_QUERY, 0, 5, 512)])
    l8, d8, c8 = ix.knn_query_padded(q, 10)
    for i, r in enumerate(probe_rows):
        assert l8[i, 0] == r and abs(d8[i, 0]) < 1e-6
    assert (np.diff(d8, axis=1) >= 0).all() and (c8 == 10).all()
    for i in range(len(q)):
        l1, d1, _ = ix.knn_query_padded(q[i:i + 1], 10)
        assert np.array_equal(l1[0], l8[i]) and np.array_equal(d1[0], d8[i])
    from oracle import c_ref
    rows = c_ref.synth_rows(R.SEED_DB, 0, n, 512)
    stored = c_ref.normalize(rows)
    want_l, want_d, _ = c_ref.knn(q, stored, None, 10, "cosine")
    for i in range(len(q)):
        msg = R.check_topk(l8[i], d8[i], q[i], stored[np.unique(np.concatenate([want_l[i], l8[i]]))],
                           np.unique(np.concatenate([want_l[i], l8[i]])), 10, "cosine", rtol=RTOL)
        assert msg is None, msg
        assert set(l8[i].tolist()) == set(want_l[i].tolist()) or msg is None


# ---------------------------------------------------------------------------------------------
# batched tensor-core path (K2 + K5 + K4): forced with set_option("path", 2)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("metric", ["l2", "ip", "cosine"])
@pytest.mark.parametrize("store,dim,shadow", [("f32", 512, 1), ("f32", 512, 0), ("f32", 100, 1), ("f16", 768, 1)])
def test_tensor_path_parity(vdb, metric, store, dim, shadow):
    """shadow=1: fp32 shard contracted through its fp16 shadow plane (kind::f16); shadow=0: the fp32 rows
    themselves (kind::tf32); fp16 shards are contracted as stored."""
    ix, raw = build(vdb, metric, 6000, dim=dim, store=store, scale=1.3)
    assert ix.get_stat("shadow") == (1 if store == "f32" else 0)
    ix.set_option("shadow", shadow)
    ix.set_option("path", 2)
    q = R.synth_rows(R.SEED_QUERY, 0, 150, dim) * np.float32(0.8)
    assert_parity(ix, raw, metric, store, q, 10)
    assert ix.get_stat("tensor_batches") >= 1


@pytest.mark.parametrize("shadow", [1, 0])
@pytest.mark.parametrize("k", [1, 16, 17, 100, 128])
def test_tensor_path_k_sizes(vdb, k, shadow):
    ix, raw = build(vdb, "l2", 9000)
    ix.set_option("shadow", shadow)
    ix.set_option("path", 2)
    q = R.synth_rows(R.SEED_QUERY, 3, 40, 512)
    assert_parity(ix, raw, "l2", "f32", q, k)


@pytest.mark.parametrize("n", [1, 100, 255, 256, 257, 1000])
def test_tensor_path_small_and_ragged(vdb, n):
    ix, raw = build(vdb, "ip", n, cap=2048)
    ix.set_option("path", 2)
    q = R.synth_rows(R.SEED_QUERY, 0, 9, 512)
    assert_parity(ix, raw, "ip", "f32", q, 10)


@pytest.mark.parametrize("n,k", [(600, 10), (1000, 10), (1500, 10), (3000, 10), (2500, 64), (4500, 100), (6000, 100)])
def test_tensor_path_small_shards_stay_on_the_tensor_path(vdb, n, k):
    """shards between one buffer and a few probes' worth of rows: the probe sees fewer chunks than the threshold
    rank (no threshold, or the largest chunk minimum) -- the buffers must still hold what the level keeps."""
    ix, raw = build(vdb, "l2", n, dim=128)
    ix.set_option("path", 2)
    q = R.synth_rows(R.SEED_QUERY, 0, 40, 128)
    assert_parity(ix, raw, "l2", "f32", q, k)
    assert ix.get_stat("tensor_batches") >= 1 and ix.get_stat("fallback_queries") == 0


def test_tensor_path_tombstones(vdb):
    ix, raw = build(vdb, "cosine", 5000)
    ix.set_option("path", 2)
    q = R.synth_rows(R.SEED_QUERY, 0, 130, 512)
    stored = R.prepare_rows(raw, "cosine")
    first, _, _ = R.knn_exact(q[:8], stored, np.arange(5000), 10, "cosine")
    dead = sorted(set(first[:, :7].reshape(-1).tolist()))
    ix.mark_deleted(dead)
    assert_parity(ix, raw, "cosine", "f32", q, 10, deleted=dead)


def test_tensor_path_equals_scan_path_bitwise(vdb):
    """K4w re-ranks with the scan kernel's summation order: all paths return identical bits."""
    ix, raw = build(vdb, "cosine", 20000)
    q = R.synth_rows(R.SEED_QUERY, 0, 300, 512)
    ix.set_option("path", 1)
    a = ix.knn_query_padded(q, 10)
    ix.set_option("path", 2)
    for shadow in (1, 0):
        ix.set_option("shadow", shadow)
        b = ix.knn_query_padded(q, 10)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])


def test_shadow_plane_follows_resize_save_load(vdb, tmp_path):
    """the fp16 shadow plane is derived data: copied by resize, rebuilt by load, extended by later adds."""
    ix, raw = build(vdb, "ip", 3000, cap=3000)
    ix.resize_index(5000)
    more = R.synth_rows(R.SEED_DB, 3000, 1500, 512)
    ix.add_items(more, np.arange(3000, 4500))
    allrows = np.concatenate([raw, more])
    q = R.synth_rows(R.SEED_QUERY, 0, 64, 512)
    ix.set_option("path", 2)
    assert_parity(ix, allrows, "ip", "f32", q, 10)
    path = str(tmp_path / "index.bin")
    ix.save_index(path)
    ix2 = vdb.Index("ip", 512)
    ix2.load_index(path, max_elements=6000)
    assert ix2.get_stat("shadow") == 1
    ix2.set_option("path", 2)
    assert_parity(ix2, allrows, "ip", "f32", q, 10)
    assert ix2.get_stat("tensor_batches") >= 1 and ix2.get_stat("fallback_queries") == 0


def test_shadow_plane_out_of_fp16_range_falls_back(vdb):
    """rows with elements beyond the fp16 range turn into inf in the shadow plane: the certificate must
    refuse and the exact scan must answer."""
    ix, raw = build(vdb, "l2", 2000, scale=3.0e6)      # elements up to ~4e5 > 65504
    ix.set_option("path", 2)
    q = R.synth_rows(R.SEED_QUERY, 0, 12, 512) * np.float32(3.0e6)
    assert_parity(ix, raw, "l2", "f32", q, 10)
    assert ix.get_stat("fallback_queries") == 12


def test_tensor_path_certificate_fallback_on_near_duplicates(vdb):
    """rows that differ below tf32 resolution cannot be ordered by the tensor pass: the coverage
    certificate must fail and the exact scan must take over -- results stay exact."""
    rng = np.random.default_rng(11)
    base = R.synth_rows(R.SEED_DB, 0, 1, 512)[0]
    rows = base[None, :] + rng.normal(0, 2e-4, size=(3000, 512)).astype(np.float32)
    ix = vdb.Index("l2", 512)
    ix.init_index(3000)
    ix.add_items(rows, np.arange(3000))
    ix.set_option("path", 2)
    q = (base[None, :] + rng.normal(0, 2e-4, size=(20, 512))).astype(np.float32)
    assert_parity(ix, rows, "l2", "f32", q, 10)
    assert ix.get_stat("fallback_queries") > 0


def test_tensor_path_large_batch_vs_scan(vdb):
    """batch 1024 over 200k rows (the bench shape, scaled): tensor path == exact scan path."""
    n = 200_000
    ix = vdb.Index("cosine", 512)
    ix.init_index(n)
    ix.add_synthetic(R.SEED_DB, 0, n)
    q = R.synth_rows(R.SEED_QUERY, 0, 1024, 512)
    ix.set_option("path", 2)
    lt, dt, ct = ix.knn_query_padded(q, 10)
    ix.set_option("path", 1)
    ls, ds, cs = ix.knn_query_padded(q[:64], 10)
    assert np.array_equal(lt[:64], ls) and np.array_equal(dt[:64], ds)
    from oracle import c_ref
    stored = c_ref.normalize(c_ref.synth_rows(R.SEED_DB, 0, n, 512))
    want_l, want_d, _ = c_ref.knn(q, stored, None, 10, "cosine")
    bad = 0
    for i in range(1024):
        if not np.array_equal(lt[i], want_l[i]):
            u = np.unique(np.concatenate([want_l[i], lt[i]]))
            msg = R.check_topk(lt[i], dt[i], q[i], stored[u], u, 10, "cosine", rtol=RTOL)
            assert msg is None, f"query {i}: {msg}"
            bad += 1
    assert bad < 16          # only distance ties within tolerance may differ
    np.testing.assert_allclose(dt, want_d, rtol=RTOL, atol=RTOL)


def test_tensor_path_batch_4096_small_blocks_select(vdb):
    """batches of >= 4096 queries take the 64-thread threshold-selection blocks: tensor path == scan path."""
    ix = vdb.Index("cosine", 512)
    ix.init_index(30000)
    ix.add_synthetic(R.SEED_DB, 0, 30000)
    q = R.synth_rows(R.SEED_QUERY, 0, 4096, 512)
    ix.set_option("path", 2)
    a = ix.knn_query_padded(q, 10)
    ix.set_option("path", 1)
    b = ix.knn_query_padded(q[-200:], 10)
    assert np.array_equal(a[0][-200:], b[0]) and np.array_equal(a[1][-200:], b[1])
    assert ix.get_stat("fallback_queries") == 0


@pytest.mark.parametrize("nq", [300, 700])
@pytest.mark.parametrize("metric,k", [("cosine", 10), ("l2", 16), ("ip", 32), ("l2", 100)])
def test_tensor_path_levels_no_fallback(vdb, metric, k, nq):
    """150k rows = 586 tiles.  700 queries (3 query blocks): probe + two threshold levels (tight rank, then k') +
    window re-rank; 300 queries: the small-batch plan (larger probe, levels growing 64x).  Random data must neither
    overflow a level buffer nor fail the certificate for any k class (k' = 32 / 64 / 256)."""
    n, dim = 150_000, 64
    ix = vdb.Index(metric, dim)
    ix.init_index(n)
    ix.add_synthetic(R.SEED_DB, 0, n)
    q = R.synth_rows(R.SEED_QUERY, 0, nq, dim)
    ix.set_option("path", 2)
    lt, dt, ct = ix.knn_query_padded(q, k)
    assert ix.get_stat("tensor_batches") >= 1 and ix.get_stat("fallback_queries") == 0
    from oracle import c_ref
    raw = c_ref.synth_rows(R.SEED_DB, 0, n, dim)
    stored = c_ref.normalize(raw) if metric == "cosine" else raw
    want_l, _, _ = c_ref.knn(q, stored, None, k, metric)
    for i in range(len(q)):
        assert ct[i] == k
        if not np.array_equal(lt[i], want_l[i]):
            u = np.unique(np.concatenate([want_l[i], lt[i]]))
            msg = R.check_topk(lt[i], dt[i], q[i], stored[u], u, k, metric, rtol=RTOL)
            assert msg is None, f"query {i}: {msg}"


def test_tensor_path_heavy_tombstones(vdb):
    """70 % of a 40k-row shard deleted (whole 32-row chunks among them): the probe must pick live rows for its
    chunk minima, the levels must drop dead survivors, results stay exact."""
    n = 40_000
    ix, raw = build(vdb, "cosine", n, dim=128)
    rng = np.random.default_rng(5)
    dead = set(np.flatnonzero(rng.random(n) < 0.6).tolist())
    for c in range(0, n // 32, 3):                     # every third chunk entirely
        dead.update(range(c * 32, c * 32 + 32))
    dead = sorted(dead)
    ix.mark_deleted(dead)
    ix.set_option("path", 2)
    q = R.synth_rows(R.SEED_QUERY, 0, 64, 128)
    assert_parity(ix, raw, "cosine", "f32", q, 10, deleted=dead)
    assert ix.get_stat("tensor_batches") >= 1 and ix.get_stat("fallback_queries") <= 2


def test_tensor_path_clustered_insertion_order(vdb):
    """rows inserted cluster by cluster (each query's neighbours sit in a few adjacent tiles): thresholds taken
    from a sample of tiles may be loose or tight, the results must not depend on it."""
    rng = np.random.default_rng(9)
    centers = rng.normal(size=(40, 128)).astype(np.float32)
    rows = np.concatenate([c[None, :] + 0.25 * rng.normal(size=(1500, 128)).astype(np.float32) for c in centers])
    n = len(rows)
    ix = vdb.Index("l2", 128)
    ix.init_index(n)
    ix.add_items(rows, np.arange(n))
    ix.set_option("path", 2)
    q = (centers[rng.integers(0, 40, size=100)] + 0.25 * rng.normal(size=(100, 128))).astype(np.float32)
    assert_parity(ix, rows, "l2", "f32", q, 10)
    assert ix.get_stat("tensor_batches") >= 1


def test_concurrent_searches_and_a_writer_on_one_index(vdb):
    """the library itself is safe for concurrent searches (one stream + workspace per call in flight) plus one
    writer (SURVEY 8b threading): 6 threads hammer scan and tensor searches while rows are appended; every
    answer must be the exact top-k of SOME prefix of the insert sequence, and the final state must be exact."""
    import threading
    n0, n1, dim, k = 4000, 6000, 512, 10
    raw = R.synth_rows(R.SEED_DB, 0, n1, dim)
    stored = R.prepare_rows(raw, "cosine")
    ix = vdb.Index("cosine", dim)
    ix.init_index(n1)
    ix.add_items(raw[:n0], np.arange(n0))
    qs = R.synth_rows(R.SEED_QUERY, 0, 48, dim)
    errors = []

    def check(l, d, q):
        # valid for a prefix: every returned id is the exact neighbour among rows [0, m) for some m >= n0:
        # cheap necessary condition -- distances are exact for the ids returned and sorted
        for i in range(len(l)):
            ids = l[i]
            want = 1.0 - stored[ids] @ R.prepare_rows(q[i:i + 1], "cosine")[0]
            if not (np.allclose(d[i], want, rtol=1e-4, atol=1e-5) and (np.diff(d[i]) >= 0).all() and (ids >= 0).all()):
                errors.append((ids.tolist(), d[i].tolist()))

    def searcher(batch):
        try:
            for _ in range(12):
                q = qs[:batch]
                l, d, c = ix.knn_query_padded(q, k)
                check(l, d, q)
        except Exception as e:          # noqa: BLE001
            errors.append(repr(e))

    def writer():
        try:
            for lo in range(n0, n1, 250):
                ix.add_items(raw[lo:lo + 250], np.arange(lo, lo + 250))
        except Exception as e:          # noqa: BLE001
            errors.append(repr(e))

    ts = [threading.Thread(target=searcher, args=(b,)) for b in (1, 3, 48, 48, 8, 20)] + [threading.Thread(target=writer)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errors, errors[:3]
    assert ix.get_current_count() == n1
    assert_parity(ix, raw, "cosine", "f32", qs, k)


@pytest.mark.parametrize("metric,store,dim", [("ip", "f32", 1408), ("l2", "f32", 1408), ("cosine", "f32", 1152),
                                              ("ip", "f16", 2816), ("l2", "f16", 1792)])
def test_widest_rows_all_positive(vdb, metric, store, dim):
    """The widest rows a shard accepts, every component positive: all `dim` products of the contraction share a
    sign, the worst case for the accumulation term (dim * 2^-23) of the tensor path's error bound.  Both paths
    must meet the parity bar, the scan in groups smaller than 8 queries where 8 no longer fit, and random data
    must stay on the tensor path (certificates hold)."""
    n, k = 12000, 10
    ix = vdb.Index(metric, dim, store_dtype=store)
    ix.init_index(n)
    raw = np.abs(R.synth_rows(R.SEED_DB, 0, n, dim))
    ix.add_items(raw, np.arange(n))
    q = np.abs(R.synth_rows(R.SEED_QUERY, 0, 24, dim)) * np.float32(0.9)
    ix.set_option("path", 2)
    assert_parity(ix, raw, metric, store, q, k)
    assert ix.get_stat("tensor_batches") == 1 and ix.get_stat("fallback_queries") == 0
    ix.set_option("path", 1)
    assert_parity(ix, raw, metric, store, q[:11], k)          # 11 queries: 8 + 3, or 4 + 4 + 3, ... by row length


def test_dims_beyond_the_ring_are_refused(vdb):
    for store, dim in (("f32", 1409), ("f16", 2817), ("f32", 8192)):
        ix = vdb.Index("l2", dim, store_dtype=store)
        with pytest.raises(RuntimeError, match="dim too large"):
            ix.init_index(100)


def test_device_call_then_host_call_share_a_workspace(vdb):
    """vdb_search_dev returns its workspace to the pool while its kernels may still run on the caller's stream; a
    host-buffer search that picks the same workspace must order itself after them (event wait), not overwrite the
    scratch in flight."""
    import torch
    n, dim, k, nq = 200_000, 512, 10, 512
    ix = vdb.Index("cosine", dim)
    ix.init_index(n)
    ix.add_synthetic(R.SEED_DB, 0, n)
    dev = torch.device("cuda", 0)
    q_np = R.synth_rows(R.SEED_QUERY, 0, nq, dim)
    q2_np = R.synth_rows(R.SEED_QUERY, 5000, nq, dim)
    want1 = ix.knn_query_padded(q_np, k)
    want2 = ix.knn_query_padded(q2_np, k)
    side = torch.cuda.Stream()
    q = torch.from_numpy(q_np).to(dev)
    ids = torch.empty((nq, k), dtype=torch.int64, device=dev)
    dd = torch.empty((nq, k), dtype=torch.float32, device=dev)
    torch.cuda.synchronize()
    for _ in range(5):
        ix.search_device(q.data_ptr(), nq, k, ids.data_ptr(), dd.data_ptr(), 0, side.cuda_stream)   # async on `side`
        got2 = ix.knn_query_padded(q2_np, k)                                                       # same workspace
        side.synchronize()
        assert np.array_equal(ids.cpu().numpy(), want1[0]) and np.array_equal(dd.cpu().numpy(), want1[1])
        assert np.array_equal(got2[0], want2[0]) and np.array_equal(got2[1], want2[1])


@pytest.mark.parametrize("k,nq", [(10, 1), (10, 2), (100, 1), (100, 2), (128, 1), (300, 1), (10, 3)])
def test_single_query_scan_merges_in_kernel(vdb, k, nq):
    """One or two queries over enough rows for a full grid: the scan kernel's last CTAs merge the per-CTA lists (one
    level at k = 10, groups + a final level at k = 100 / 128); k = 300 and three queries keep the merge kernel.  Same
    answers either way, and one launch per search when the merge is fused."""
    n, dim = 60_000, 512
    ix = vdb.Index("l2", dim)
    ix.init_index(n)
    ix.add_synthetic(R.SEED_DB, 0, n)
    ix.mark_deleted([11, 40_000])
    stored = R.synth_rows(R.SEED_DB, 0, n, dim)
    q = R.synth_rows(R.SEED_QUERY, 3, nq, dim)
    l0 = vdb.launch_count()
    got_l, got_d, cnt = ix.knn_query_padded(q, k)
    launches = vdb.launch_count() - l0
    assert launches == (1 if (nq <= 2 and k <= 128) else 2)
    for i in range(nq):
        assert cnt[i] == k
        msg = R.check_topk(got_l[i], got_d[i], q[i], stored, np.arange(n), k, "l2", deleted=[11, 40_000], rtol=RTOL)
        assert msg is None, f"query {i}: {msg}"
    again = ix.knn_query_padded(q, k)                              # the arrival counters were left at zero
    assert np.array_equal(again[0], got_l) and np.array_equal(again[1], got_d)


def test_device_side_fallback_with_arbitrary_labels(vdb):
    """A flood of near-duplicates fails the tensor path's certificates: the flagged queries are re-searched exactly by
    a kernel of the same enqueue (no host round trip); labels that are not base + row go through the label array."""
    dim, n = 512, 3000
    base = R.synth_rows(R.SEED_DB, 0, 1, dim)[0]
    dup = np.tile(base, (n, 1)) + np.random.default_rng(1).normal(0, 1e-6, (n, dim)).astype(np.float32)
    labels = np.arange(n)[::-1].copy() * 3 + 7
    ix = vdb.Index("cosine", dim)
    ix.init_index(n)
    ix.add_items(dup, labels)
    ix.set_option("path", 2)
    q = R.synth_rows(R.SEED_QUERY, 0, 12, dim)
    got_l, got_d, cnt = ix.knn_query_padded(q, 10)
    stored = R.prepare_rows(dup, "cosine")
    for i in range(len(q)):
        msg = R.check_topk(got_l[i], got_d[i], q[i], stored, labels, 10, "cosine", rtol=RTOL)
        assert msg is None, f"query {i}: {msg}"
    assert ix.get_stat("fallback_queries") > 0 and ix.get_stat("tensor_batches") == 1


@pytest.mark.parametrize("pinned", [True, False])
def test_submit_collect_two_batches_in_flight(vdb, pinned):
    """vdb_search_submit / vdb_search_collect: several tickets open at once (tensor path and scan path mixed),
    collected in order, return exactly what the blocking call returns; an empty batch gives an empty ticket."""
    ix, raw = build(vdb, "cosine", 30_000)
    k = 10
    batches = [R.synth_rows(R.SEED_QUERY, 1000 * i, nq, 512) for i, nq in enumerate((64, 1, 300, 2, 64))]
    want = [tuple(a.copy() for a in ix.knn_query_padded(q, k)) for q in batches]
    alloc = vdb.pinned_empty if pinned else (lambda shape, dt: np.empty(shape, dtype=dt))
    for depth in (2, 3):
        pending, got = [], []
        for q in batches:
            qh = alloc(q.shape, np.float32)
            qh[:] = q
            out = (alloc((len(q), k), np.int64), alloc((len(q), k), np.float32), alloc((len(q),), np.int32))
            pending.append(ix.submit_query(qh, k, out=out))
            if len(pending) == depth:
                got.append(tuple(a.copy() for a in ix.collect_query(pending.pop(0))))
        while pending:
            got.append(tuple(a.copy() for a in ix.collect_query(pending.pop(0))))
        for (wl, wd, wc), (gl, gd, gc) in zip(want, got):
            assert (wl == gl).all() and (wd == gd).all() and (wc == gc).all()
    l, d, c = ix.collect_query(ix.submit_query(np.zeros((0, 512), np.float32), k))
    assert l.shape == (0, k) and c.shape == (0,)


# ---------------------------------------------------------------------------------------------
# one or two queries on an fp32 shard with an fp16 shadow plane: K1 over the shadow + exact re-rank
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("metric", ["l2", "ip", "cosine"])
@pytest.mark.parametrize("nq,k", [(1, 10), (2, 10), (1, 1), (2, 32), (1, 17)])
def test_shadow_scan_parity_and_bitwise_equal_to_the_fp32_scan(vdb, metric, nq, k):
    """The shadow-plane scan returns exactly what the fp32 scan returns (ids, distances bit for bit), tombstones and
    labels that are not base + row included, and matches the oracle."""
    n, dim = 20_000, 512
    raw = R.synth_rows(R.SEED_DB, 0, n, dim) * np.float32(1.3)
    labels = (np.arange(n)[::-1] * 2 + 5).astype(np.int64)
    ix = vdb.Index(metric, dim)
    ix.init_index(n)
    ix.add_items(raw, labels)
    dead = labels[[3, 500, 7777, 19_999]]
    ix.mark_deleted(dead)
    q = R.synth_rows(R.SEED_QUERY, 31, nq, dim) * np.float32(0.8)
    ix.set_option("shadow_scan_rows", 1 << 40)            # off: the fp32 scan
    want = tuple(a.copy() for a in ix.knn_query_padded(q, k))
    assert ix.get_stat("shadow_scans") == 0
    ix.set_option("shadow_scan_rows", 1000)
    ix.set_option("shadow_scan_nq", 2)                    # default: single queries only (two are faster on the tensor path)
    got = ix.knn_query_padded(q, k)
    assert ix.get_stat("shadow_scans") == 1 and ix.get_stat("fallback_queries") == 0
    assert (got[0] == want[0]).all() and (got[1] == want[1]).all() and (got[2] == want[2]).all()
    stored = R.prepare_rows(raw, metric)
    for i in range(nq):
        msg = R.check_topk(got[0][i], got[1][i], q[i], stored, labels, k, metric, deleted=list(dead), rtol=RTOL)
        assert msg is None, f"query {i}: {msg}"


def test_shadow_scan_near_duplicates_fall_back_exactly(vdb):
    """rows that differ below fp16 resolution: the certificate of the shadow scan fails, the flagged query is
    re-searched exactly on the device; results equal the fp32 scan's."""
    rng = np.random.default_rng(5)
    base = R.synth_rows(R.SEED_DB, 0, 1, 512)[0]
    rows = base[None, :] + rng.normal(0, 1e-5, size=(4000, 512)).astype(np.float32)
    for metric in ("l2", "cosine"):
        ix = vdb.Index(metric, 512)
        ix.init_index(4000)
        ix.add_items(rows, np.arange(4000))
        q = (base[None, :] + rng.normal(0, 1e-5, size=(2, 512))).astype(np.float32)
        ix.set_option("shadow_scan_rows", 1 << 40)
        want = tuple(a.copy() for a in ix.knn_query_padded(q, 10))
        ix.set_option("shadow_scan_rows", 1000)
        ix.set_option("shadow_scan_nq", 2)
        got = ix.knn_query_padded(q, 10)
        assert ix.get_stat("shadow_scans") == 1 and ix.get_stat("fallback_queries") > 0
        assert (got[0] == want[0]).all() and (got[1] == want[1]).all()


def test_shadow_scan_short_results_and_wide_norm_spread(vdb):
    """fewer live rows than k' (everything is a candidate, no certificate needed); rows whose norms differ by 100x
    with an L2 query far from all of them (the direct-form error terms of the bound)."""
    dim = 512
    raw = R.synth_rows(R.SEED_DB, 0, 3000, dim)
    raw[::7] *= np.float32(100.0)
    ix = vdb.Index("l2", dim)
    ix.init_index(3000)
    ix.add_items(raw, np.arange(3000))
    ix.set_option("shadow_scan_rows", 1000)
    ix.set_option("shadow_scan_nq", 2)
    q = R.synth_rows(R.SEED_QUERY, 3, 2, dim) * np.float32(0.01)
    assert_parity(ix, raw, "l2", "f32", q, 10)
    assert_parity(ix, raw, "l2", "f32", q * np.float32(5000.0), 10)
    assert ix.get_stat("shadow_scans") == 2
    ix.mark_deleted(np.arange(20, 3000))
    assert_parity(ix, raw, "l2", "f32", q[:1], 10, deleted=list(range(20, 3000)))     # 20 live rows < k' = 32
