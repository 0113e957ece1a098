"""GPU: the datanode handler with the real CUDA index behind it == the oracle's datanode model, including
recovery from checkpoint + WAL, and the in-process coordinator over several shards on one GPU."""
import numpy as np
import pytest

from oracle import cpu_ref as R

pytestmark = pytest.mark.gpu
DIM = 512


def vecs(n, seed=R.SEED_DB, start=0):
    return R.synth_rows(seed, start, n, DIM)


def test_gpu_handler_put_delete_search_recover(vdb, tmp_path):
    h = vdb.GpuVectorNodeHandler("node_1", storage_root=str(tmp_path), space="cosine", dim=DIM, max_elements=4096,
                                 checkpoint_every=300, fsync=False)
    model = R.DatanodeModel(dim=DIM, metric="cosine")
    rows = vecs(700)
    for i in range(400):
        assert h.put(vdb.VectorData(key=f"k{i}", vector=rows[i].tolist(), metadata={"i": str(i)})).success
        model.put(f"k{i}", rows[i].tolist())
    items = [vdb.VectorData(key=f"k{i}", vector=rows[i].tolist(), metadata={}) for i in range(350, 700)]   # 50 overwrites
    assert h.put_batch(items).success
    for i in range(350, 700):
        model.put(f"k{i}", rows[i].tolist())
    for i in range(0, 60, 3):
        h.delete(f"k{i}"); model.delete(f"k{i}")
    qs = R.synth_rows(R.SEED_QUERY, 0, 6, DIM)
    qs[0] = rows[10]                                   # exact hit on a live row
    for q in qs:
        got = h.search(vdb.SearchRequest(query_vector=q.tolist(), top_k=10)).search_result
        ok, keys, scores = R.datanode_search_exact_live(model, q.tolist(), 10)
        assert got.keys == keys
        np.testing.assert_allclose(got.scores, scores, rtol=1e-5, atol=1e-6)
    r = h.search(vdb.SearchRequest(query_vector=rows[10].tolist(), top_k=1)).search_result
    assert r.keys == ["k10"] and abs(r.scores[0]) < 1e-6 and r.vectors[0].metadata == {"i": "10"}
    ks, ss = h.search_batch(qs, 10)                    # tensor-core path for nq > 8? (6 here: scan) -- same answers
    assert ks[1] == h.search(vdb.SearchRequest(query_vector=qs[1].tolist(), top_k=10)).search_result.keys
    want = [h.search(vdb.SearchRequest(query_vector=q.tolist(), top_k=10)).search_result for q in qs]
    # crash + restart on the same directory: newest checkpoint + incremental WAL replay
    h2 = vdb.GpuVectorNodeHandler("node_1", storage_root=str(tmp_path), space="cosine", dim=DIM, max_elements=4096,
                                  checkpoint_every=300, fsync=False)
    for q, w in zip(qs, want):
        got = h2.search(vdb.SearchRequest(query_vector=q.tolist(), top_k=10)).search_result
        assert got.keys == w.keys
        np.testing.assert_allclose(got.scores, w.scores, rtol=1e-6, atol=1e-7)


def test_local_coordinator_over_gpu_shards(vdb, tmp_path):
    nodes = {f"n{i}": vdb.GpuVectorNodeHandler(f"n{i}", storage_root=str(tmp_path), space="l2", dim=DIM, max_elements=1024,
                                               checkpoint_every=0, fsync=False) for i in range(4)}
    coord = vdb.LocalCoordinator(nodes)
    model = R.DatanodeModel(dim=DIM, metric="l2")
    rows = vecs(600)
    for i in range(600):
        coord.put(vdb.VectorData(key=f"img_{i}", vector=rows[i].tolist(), metadata={}))
        model.put(f"img_{i}", rows[i].tolist())
    counts = [h.hnsw_index.get_current_count() for h in nodes.values()]
    assert sum(counts) == 600 and min(counts) > 100                       # md5 routing spreads the keys
    for q in R.synth_rows(R.SEED_QUERY, 0, 4, DIM):
        got = coord.search(vdb.SearchRequest(query_vector=q.tolist(), top_k=10)).search_result
        ok, keys, scores = R.datanode_search_exact_live(model, q.tolist(), 10)
        assert got.keys == keys
        np.testing.assert_allclose(got.scores, scores, rtol=1e-5, atol=1e-6)


def test_concurrent_searches_and_writer(vdb):
    """TThreadPoolServer runs 5 workers against one handler (datanode/server.py:25-28): concurrent searches
    with a concurrent writer must stay correct."""
    import threading
    ix = vdb.Index("ip", DIM)
    ix.init_index(20000)
    ix.add_synthetic(R.SEED_DB, 0, 8000)
    q = R.synth_rows(R.SEED_QUERY, 0, 4, DIM)
    base = ix.knn_query_padded(q, 10)
    errs = []

    def searcher():
        try:
            for _ in range(30):
                l, d, c = ix.knn_query_padded(q, 10)
                # rows only get appended: a result can only improve, and must stay sorted
                assert (np.diff(d, axis=1) >= 0).all() and (d[:, 0] <= base[1][:, 0] + 1e-6).all()
        except Exception as e:      # noqa
            errs.append(e)

    def writer():
        try:
            for j in range(10):
                ix.add_synthetic(R.SEED_DB, 8000 + j * 500, 500)
        except Exception as e:      # noqa
            errs.append(e)

    ts = [threading.Thread(target=searcher) for _ in range(5)] + [threading.Thread(target=writer)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs, errs
    assert ix.get_current_count() == 13000
    rows = R.synth_rows(R.SEED_DB, 0, 13000, DIM)
    l, d, c = ix.knn_query_padded(q, 10)
    for i in range(4):
        assert R.check_topk(l[i], d[i], q[i], rows, np.arange(13000), 10, "ip", rtol=1e-5) is None


def test_micro_batched_handler_on_the_gpu(vdb, tmp_path):
    """8 worker threads of single-query searches coalesced into tensor-path batches: same keys and scores (bit for
    bit: the re-rank uses the scan kernel's summation order) as one query at a time, with a writer running."""
    import threading
    n = 3000
    rows = vecs(n)
    hs = [vdb.GpuVectorNodeHandler(f"node_{i}", storage_root=str(tmp_path), space="cosine", dim=DIM, max_elements=8192,
                                   checkpoint_every=0, fsync=False, micro_batch_wait_s=w) for i, w in enumerate((None, 1e-3))]
    items = [vdb.VectorData(key=f"k{i}", vector=rows[i].tolist(), metadata={"i": str(i)}) for i in range(n)]
    for h in hs:
        assert h.put_batch(items).success
    plain, fast = hs
    qs = R.synth_rows(R.SEED_QUERY, 0, 24, DIM)
    reqs = [vdb.SearchRequest(query_vector=q.tolist(), top_k=10) for q in qs]
    want = [plain.search(r).search_result for r in reqs]
    got, errs = {}, []

    def worker(t):
        try:
            for j, r in enumerate(reqs):
                got[(t, j)] = fast.search(r)
        except Exception as e:          # noqa
            errs.append(e)

    def writer():                       # appended rows are far from every query: answers must not change
        try:
            far = -qs.sum(axis=0)
            for j in range(20):
                assert fast.put(vdb.VectorData(key=f"far{j}", vector=far.tolist(), metadata={})).success
        except Exception as e:          # noqa
            errs.append(e)

    ts = [threading.Thread(target=worker, args=(t,)) for t in range(8)] + [threading.Thread(target=writer)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs, errs
    for (t, j), resp in got.items():
        assert resp.success and resp.search_result.keys == want[j].keys
        assert resp.search_result.scores == want[j].scores
    assert fast._batcher.requests == 8 * 24 and fast._batcher.batches < 8 * 24
    assert fast.hnsw_index.get_stat("tensor_batches") >= 1


def test_coordinator_search_batch_gpu_merge_equals_host_merge(vdb, tmp_path):
    """LocalCoordinator.search_batch: per-node [nq, k] arrays merged by the GPU merge kernel (vdb_merge_topk), ids
    mapped to keys once -- against the reference's merge rule in Python (coordinator/handler.py:200-225) and the
    single-node model."""
    nodes = {f"n{i}": vdb.GpuVectorNodeHandler(f"n{i}", storage_root=str(tmp_path), space="cosine", dim=DIM,
                                               max_elements=2048, checkpoint_every=0, fsync=False) for i in range(4)}
    coord = vdb.LocalCoordinator(nodes)
    model = R.DatanodeModel(dim=DIM, metric="cosine")
    rows = vecs(1500)
    keys = [f"img_{i}" for i in range(1500)]
    by_node = {}
    for i, key in enumerate(keys):
        by_node.setdefault(coord.mapping[vdb.get_shard_id(key, coord.shard_count)]["master"], []).append(i)
        model.put(key, rows[i].tolist())
    for node_id, idx in by_node.items():
        assert nodes[node_id].put_arrays([keys[i] for i in idx], rows[idx]).success
    for i in (3, 77, 500):                                   # tombstones on several shards
        coord.delete(keys[i]); model.delete(keys[i])
    qs = R.synth_rows(R.SEED_QUERY, 0, 40, DIM)
    qs[0] = rows[9]
    for k in (1, 10, 50):
        gk, gs = coord.search_batch(qs, k)                    # GPU merge
        hk, hs = coord.search_batch(qs, k, gpu_merge=False)   # the reference's merge in Python
        assert gk == hk
        for a, b in zip(gs, hs):
            np.testing.assert_allclose(a, b, rtol=0, atol=0)
        for r in (0, 7, 39):
            ok, mk, ms = R.datanode_search_exact_live(model, qs[r].tolist(), k)
            assert gk[r] == mk
            np.testing.assert_allclose(gs[r], ms, rtol=1e-5, atol=1e-6)
    assert gk[0][0] == "img_9"
    # a node that fails is skipped (coordinator/handler.py:198-199)
    class Down:
        device = 0
        def search_ids(self, q, k): raise ConnectionError("down")
        def keys_of(self, ids): return []
    flaky = vdb.LocalCoordinator({**nodes, "zz": Down()}, shard_count=4)
    fk, fs = flaky.search_batch(qs[:3], 10)
    assert fk == gk_ref(coord, qs[:3], 10)


def gk_ref(coord, qs, k):
    return coord.search_batch(qs, k, gpu_merge=False)[0]


def test_shard_grows_in_place_while_searches_run(vdb):
    """resize_index maps more memory behind the same base pointers (no copy, no exclusive lock): searches that run
    during the growth and the appends stay correct; growing past the address reservation (8x) re-bases without
    copying; the content survives both."""
    import threading
    ix = vdb.Index("l2", DIM)
    ix.init_index(1000)
    raw = R.synth_rows(R.SEED_DB, 0, 1000, DIM)
    ix.add_items(raw, np.arange(1000))
    q = R.synth_rows(R.SEED_QUERY, 0, 3, DIM)
    base = ix.knn_query_padded(q, 10)
    errs, stop = [], threading.Event()

    def searcher():
        try:
            while not stop.is_set():
                l, d, c = ix.knn_query_padded(q, 10)
                assert (np.diff(d, axis=1) >= 0).all() and (d[:, 0] <= base[1][:, 0] + 1e-6).all() and (c == 10).all()
        except Exception as e:      # noqa
            errs.append(e)

    ts = [threading.Thread(target=searcher) for _ in range(3)]
    [t.start() for t in ts]
    try:
        with pytest.raises(RuntimeError):                     # full, as hnswlib
            ix.add_items(raw[:1], [5000])
        for cap in (3000, 9000, 1_200_000, 9_000_000):        # the last one is beyond the reservation: re-base
            ix.resize_index(cap)
            assert ix.get_max_elements() == cap
            n0 = ix.get_current_count()
            ix.add_synthetic(R.SEED_DB, n0, 1500)
    finally:
        stop.set()
        [t.join() for t in ts]
    assert not errs, errs
    n = ix.get_current_count()
    assert n == 7000
    stored = R.synth_rows(R.SEED_DB, 0, n, DIM)
    got_l, got_d, cnt = ix.knn_query_padded(q, 10)
    for i in range(len(q)):
        msg = R.check_topk(got_l[i], got_d[i], q[i], stored, np.arange(n), 10, "l2", rtol=1e-5)
        assert msg is None, msg
    ix.resize_index(8000)                                     # shrinking the limit (>= count) is allowed
    assert ix.get_max_elements() == 8000
    with pytest.raises(RuntimeError):
        ix.resize_index(10)


def test_handler_bulk_load_and_recovery(vdb, tmp_path):
    """put_arrays: one add_items + one key-store append + one WAL group commit per batch (vectors logged by reference),
    the shard grows in place past max_elements, and a restart restores keys, vectors and metadata."""
    h = vdb.GpuVectorNodeHandler("bulk", storage_root=str(tmp_path), space="cosine", dim=DIM, max_elements=1000,
                                 checkpoint_every=2500, fsync=False)
    rows = vecs(4000)
    keys = [f"b{i}" for i in range(4000)]
    for lo in range(0, 4000, 1000):
        assert h.put_arrays(keys[lo:lo + 1000], rows[lo:lo + 1000], [{"i": str(i)} if i % 7 == 0 else None
                                                                     for i in range(lo, lo + 1000)]).success
    assert h.hnsw_index.get_current_count() == 4000 and h.hnsw_index.get_max_elements() >= 4000
    extra = vecs(3, start=100_000)
    assert h.put_arrays(["b5", "b5", "new"], extra).success                          # overwrite + in-batch duplicate
    qs = np.stack([extra[1], rows[1234], extra[2], rows[5], extra[0]])
    want = h.search_batch(qs, 5)
    assert [w[0] for w in want[0][:3]] == ["b5", "b1234", "new"]
    assert "b5" not in want[0][3][:1] and want[1][3][0] > 1e-3                       # the old b5 row is tombstoned
    assert want[1][4][0] > 1e-3                                                      # and so is the in-batch duplicate
    g = h.get("b7")
    assert g.vector_data.metadata == {"i": "7"} and np.allclose(g.vector_data.vector, rows[7])
    h.store.close()                                           # crash: no final checkpoint
    h2 = vdb.GpuVectorNodeHandler("bulk", storage_root=str(tmp_path), space="cosine", dim=DIM, max_elements=1000,
                                  checkpoint_every=2500, fsync=False)
    got = h2.search_batch(qs, 5)
    assert got[0] == want[0]
    np.testing.assert_allclose(np.array(got[1]), np.array(want[1]), rtol=1e-6, atol=1e-7)
    g = h2.get("b7")
    assert g.vector_data.metadata == {"i": "7"} and np.allclose(g.vector_data.vector, rows[7])
    assert np.allclose(h2.get("b5").vector_data.vector, extra[1]) and len(h2.store) == 4001


def test_append_only_shard_image(vdb, tmp_path):
    """vdb_save_image / vdb_load_image: the on-disk image grows by the rows added since the last save, every meta
    file stays loadable (its prefix never changes), a crashed save (files longer than the newest meta) and a restart
    from an older meta are cut back before the next append."""
    img = str(tmp_path / "img")
    n1, n2 = 3000, 5000
    raw = vecs(n2)
    ix = vdb.Index("cosine", DIM)
    ix.init_index(n2)
    ix.add_items(raw[:n1], np.arange(n1))
    ix.mark_deleted([5, 7])
    ix.save_image(img, str(tmp_path / "meta1.bin"))
    size1 = (tmp_path / "img" / "rows.bin").stat().st_size
    ix.add_items(raw[n1:], np.arange(n1, n2))
    ix.mark_deleted([4000])
    ix.save_image(img, str(tmp_path / "meta2.bin"))
    assert (tmp_path / "img" / "rows.bin").stat().st_size == size1 * n2 // n1        # appended, not rewritten
    assert (tmp_path / "meta2.bin").stat().st_size < 4096                            # a checkpoint is a few hundred bytes
    q = R.synth_rows(R.SEED_QUERY, 0, 20, DIM)
    want2 = ix.knn_query_padded(q, 10)
    stored = R.prepare_rows(raw, "cosine")
    # newest meta: the whole shard
    b = vdb.Index("cosine", DIM)
    b.load_image(img, str(tmp_path / "meta2.bin"), max_elements=n2 + 100)
    got = b.knn_query_padded(q, 10)
    assert b.get_current_count() == n2 and np.array_equal(got[0], want2[0]) and np.array_equal(got[1], want2[1])
    # older meta: the prefix, with the tombstones of that moment
    a = vdb.Index("cosine", DIM)
    a.load_image(img, str(tmp_path / "meta1.bin"), max_elements=n2)
    assert a.get_current_count() == n1 and a.get_live_count() == n1 - 2
    l, d, c = a.knn_query_padded(q[:4], 10)
    for i in range(4):
        msg = R.check_topk(l[i], d[i], q[i], stored[:n1], np.arange(n1), 10, "cosine", deleted=[5, 7], rtol=1e-5)
        assert msg is None, msg
    # continue from the OLDER state with different rows: the stale tail of the image is cut off first
    other = vecs(500, start=900_000)
    a.add_items(other, np.arange(n1, n1 + 500))
    a.save_image(img, str(tmp_path / "meta3.bin"))
    assert (tmp_path / "img" / "rows.bin").stat().st_size == size1 * (n1 + 500) // n1
    c3 = vdb.Index("cosine", DIM)
    c3.load_image(img, str(tmp_path / "meta3.bin"))
    l, d, c = c3.knn_query_padded(other[:3], 1)
    assert l[:, 0].tolist() == [n1, n1 + 1, n1 + 2] and np.all(np.abs(d[:, 0]) < 1e-6)
    with pytest.raises(RuntimeError):
        vdb.Index("cosine", DIM).load_image(img, str(tmp_path / "meta2.bin"))       # names rows the image no longer has
    assert vdb.Index.is_image_meta(str(tmp_path / "meta3.bin")) and not vdb.Index.is_image_meta(str(tmp_path / "img" / "rows.bin"))


def test_handler_checkpoints_are_incremental(vdb, tmp_path):
    """checkpoint_every = 1000 with 1000-row batches: every batch takes a checkpoint (the reference's cadence,
    handler.py:316-317); with the append-only image each costs the new rows only, and a restart from the newest one
    (plus the WAL tail) restores the state."""
    h = vdb.GpuVectorNodeHandler("inc", storage_root=str(tmp_path), space="l2", dim=DIM, max_elements=2000,
                                 checkpoint_every=1000, fsync=False)
    rows = vecs(6500)
    for lo in range(0, 6000, 1000):
        assert h.put_arrays([f"k{i}" for i in range(lo, lo + 1000)], rows[lo:lo + 1000]).success
    assert h.put_arrays([f"k{i}" for i in range(6000, 6500)], rows[6000:]).success   # WAL tail, no checkpoint
    h.delete("k10")
    cps = sorted(d for d in (tmp_path / "inc" / "checkpoint").iterdir())
    assert len(cps) == 2 and all((c / "index.bin").stat().st_size < 4096 for c in cps)
    assert (tmp_path / "inc" / "hnsw_index" / "rows.bin").stat().st_size == 6000 * DIM * 4
    qs = rows[[10, 3333, 6400]]
    want = h.search_batch(qs, 3)
    h.store.close()
    h2 = vdb.GpuVectorNodeHandler("inc", storage_root=str(tmp_path), space="l2", dim=DIM, max_elements=2000,
                                  checkpoint_every=1000, fsync=False)
    got = h2.search_batch(qs, 3)
    assert got[0] == want[0] and "k10" not in got[0][0] and got[0][1][0] == "k3333" and got[0][2][0] == "k6400"
    np.testing.assert_allclose(np.array(got[1]), np.array(want[1]), rtol=1e-6, atol=1e-7)
    assert len(h2.store) == 6499
