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
