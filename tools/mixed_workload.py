#!/usr/bin/env python
"""BASELINE config 5 (mixed insert + search through the WAL on 4 shards, cosine top-10, batch 256), scaled.

Four `GpuVectorNodeHandler`s (one shard each, md5 routing through `LocalCoordinator`'s rule) on one GPU.  The
database grows from --start to --end rows in --insert-batch puts (`put_batch`: one add_items + one WAL group
commit per shard); after every growth step a batch of 256 queries is searched on every shard
(`search_batch`, the tensor path) and the per-shard lists are merged with the GPU merge kernel.  Reports search
queries/s against N, insert rows/s and WAL bytes/s, then checks the final state against the CPU oracle and a
cold restart (checkpoint + WAL replay) against the live state.

Handler level: `put_arrays` (one add_items + one key-store append + one WAL group commit per shard, vectors logged
by reference), shards that start small and grow in place, `LocalCoordinator.search_batch` (GPU merge), a searcher
thread that keeps searching WHILE the inserts run (p50 latency against the idle p50), a sampled oracle check of
the final state and a checkpoint + WAL-tail restart of one shard.  Default 1/20 scale; `--start 1000000 --end
5000000 --insert-batch 200000` is BASELINE config 5 at full size (needs ~12 GB of disk for the raw-vector files).
`--level index` runs the same growth/search schedule on bare `Index` shards."""
import argparse, json, os, shutil, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import dvdb_b200 as vdb
from oracle import c_ref, cpu_ref as R

ap = argparse.ArgumentParser()
ap.add_argument("--start", type=int, default=50_000)
ap.add_argument("--end", type=int, default=250_000)
ap.add_argument("--insert-batch", type=int, default=50_000)
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--shards", type=int, default=4)
ap.add_argument("--dim", type=int, default=512)
ap.add_argument("--k", type=int, default=10)
ap.add_argument("--level", default="handler", choices=["handler", "index"])
ap.add_argument("--searches-per-step", type=int, default=5)
ap.add_argument("--tmp", default="", help="directory for the datanode storage roots (default: the system temp dir)")
ap.add_argument("--fsync", action="store_true", help="fsync the WAL / key store on every commit")
a = ap.parse_args()

keys_of = lambda lo, hi: [f"k{r:09d}" for r in range(lo, hi)]
queries = R.synth_rows(R.SEED_QUERY, 0, a.batch, a.dim)
report = {"config": vars(a), "steps": []}


def dir_bytes(path):
    return sum(os.path.getsize(os.path.join(d, f)) for d, _, fs in os.walk(path) for f in fs)


def merged_search(search_fns, nq):
    """every shard answers the whole batch; GPU merge by (distance, id) over the shard axis"""
    ids = np.full((len(search_fns), nq, a.k), -1, np.int64)
    dd = np.full((len(search_fns), nq, a.k), np.inf, np.float32)
    for s, fn in enumerate(search_fns):
        l, d, c = fn(queries, a.k)
        ids[s], dd[s] = l, d
    return vdb.merge_topk(dd, ids, a.k)


if a.level == "handler":
    import importlib.util, statistics, threading
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec); sys.modules["bench_mod"] = bench; spec.loader.exec_module(bench)
    c_ref.set_threads(c_ref.host_cores())
    root = tempfile.mkdtemp(prefix="vdb_mixed_", dir=a.tmp or None)
    try:
        handlers = {f"node_{s}": vdb.GpuVectorNodeHandler(f"node_{s}", root, space="cosine", dim=a.dim,
                                                          max_elements=max(a.start // a.shards, 1024), checkpoint_every=0,
                                                          fsync=a.fsync)
                    for s in range(a.shards)}
        coord = vdb.LocalCoordinator(handlers, shard_count=a.shards)
        node_ids = coord.node_ids

        def put_rows(lo, hi):
            """one put_arrays per shard: add_items + key-store append + WAL group commit (vectors by reference)"""
            rows = c_ref.synth_rows(R.SEED_DB, lo, hi - lo, a.dim)
            keys = keys_of(lo, hi)
            owner = np.array([node_ids.index(coord.mapping[vdb.get_shard_id(k, a.shards)]["master"]) for k in keys])
            t = time.perf_counter()
            for s, n in enumerate(node_ids):
                sel = np.nonzero(owner == s)[0]
                resp = handlers[n].put_arrays([keys[i] for i in sel], rows[sel])
                assert resp.success, resp.message
            return time.perf_counter() - t

        lat_busy, lat_idle, stop = [], [], threading.Event()

        def searcher(sink):
            while not stop.is_set():
                t = time.perf_counter()
                coord.search_batch(queries, a.k)
                sink.append(time.perf_counter() - t)

        n = 0
        wal0 = 0
        while n < a.end:
            hi = min(a.end, n + (a.start if n == 0 else a.insert_batch))
            sink = []
            stop.clear()
            th = threading.Thread(target=searcher, args=(sink,)) if n else None    # searches DURING the inserts
            if th: th.start()
            dt_ins = put_rows(n, hi)
            stop.set()
            if th: th.join()
            wal1 = sum(dir_bytes(handlers[x].wal_dir) for x in node_ids)
            coord.search_batch(queries, a.k)
            idle = []
            for _ in range(a.searches_per_step):
                t = time.perf_counter()
                res = coord.search_batch(queries, a.k)
                idle.append(time.perf_counter() - t)
            dt_s = statistics.median(idle)
            step = {"rows": hi, "insert_rows_per_s": (hi - n) / dt_ins, "wal_MB_per_s": (wal1 - wal0) / dt_ins / 1e6,
                    "search_qps": a.batch / dt_s, "search_ms_p50_idle": 1e3 * dt_s}
            if sink:
                step["search_ms_p50_during_inserts"] = 1e3 * statistics.median(sink)
                step["searches_during_inserts"] = len(sink)
            report["steps"].append(step)
            n, wal0 = hi, wal1
        report["capacity_per_shard"] = [handlers[x].hnsw_index.get_max_elements() for x in node_ids]
        # final state against the oracle: sampled queries, exact CPU scan of the whole set (chunked)
        q_idx = np.unique(np.linspace(0, a.batch - 1, min(64, a.batch)).astype(np.int64))
        wl = bench.Workload("config5", a.end, a.dim, a.k, "cosine", "f32", a.batch)
        want_l, want_d, info = bench.oracle_topk(wl, q_idx, budget_s=600.0)
        got_keys, got_scores = res
        bad = 0
        for j, qi in enumerate(q_idx):
            got = [int(k[1:]) for k in got_keys[qi]]
            if got != want_l[j].tolist():
                bad += 1
            assert np.allclose(np.array(got_scores[qi], np.float32), want_d[j], rtol=1e-5, atol=1e-6), f"query {qi}"
        report["oracle_parity"] = f"{len(q_idx)} sampled queries vs exact CPU scan of {a.end} rows: {bad} with a different id order (distance ties within 1e-5), all distances within 1e-5"
        # cold restart of shard 0: checkpoint + WAL tail
        h0 = handlers[node_ids[0]]
        live = len(h0.store)
        t = time.perf_counter(); h0.save_checkpoint(); report["checkpoint_s"] = time.perf_counter() - t
        extra = c_ref.synth_rows(R.SEED_DB, a.end, 1000, a.dim)
        h0.put_arrays([f"tail{i}" for i in range(1000)], extra)                    # after the checkpoint: replayed from the WAL
        k0, s0 = h0.search_batch(queries[:16], a.k)
        h0.store.close(); h0.hnsw_index.close()
        t = time.perf_counter()
        h0b = vdb.GpuVectorNodeHandler(node_ids[0], root, space="cosine", dim=a.dim, max_elements=1024, checkpoint_every=0,
                                       fsync=a.fsync)
        report["restart_s"] = time.perf_counter() - t
        assert len(h0b.store) == live + 1000, "restart lost or invented keys"
        k1, s1 = h0b.search_batch(queries[:16], a.k)
        assert k1 == k0 and np.allclose(np.array(s1), np.array(s0), rtol=1e-6, atol=1e-7)
        report["restart"] = f"shard 0: {live} keys from the checkpoint + 1000 from the WAL tail, searches identical"
        report["disk_bytes"] = dir_bytes(root)
    finally:
        shutil.rmtree(root, ignore_errors=True)
else:
    shards = []
    for s in range(a.shards):
        ix = vdb.Index("cosine", a.dim)
        ix.init_index(a.end)              # md5 routing is not exactly even: room for the whole set
        shards.append(ix)

    def put_rows(lo, hi):
        rows = c_ref.synth_rows(R.SEED_DB, lo, hi - lo, a.dim)
        sid = np.array([vdb.get_shard_id(k, a.shards) for k in keys_of(lo, hi)])
        t = time.perf_counter()
        for s, ix in enumerate(shards):
            sel = np.nonzero(sid == s)[0]
            ix.add_items(rows[sel], sel + lo)
        return time.perf_counter() - t

    n = 0
    while n < a.end:
        hi = min(a.end, n + (a.start if n == 0 else a.insert_batch))
        dt_ins = put_rows(n, hi)
        fns = [ix.knn_query_padded for ix in shards]
        merged_search(fns, a.batch)
        t = time.perf_counter()
        for _ in range(a.searches_per_step):
            od, oi = merged_search(fns, a.batch)
        dt_s = (time.perf_counter() - t) / a.searches_per_step
        report["steps"].append({"rows": hi, "insert_rows_per_s": (hi - n) / dt_ins, "search_qps": a.batch / dt_s, "search_ms": 1e3 * dt_s})
        n = hi
    if a.end <= 2_000_000:
        stored = c_ref.normalize(c_ref.synth_rows(R.SEED_DB, 0, a.end, a.dim))
        want_l, want_d, _ = c_ref.knn(queries, stored, None, a.k, "cosine")
        bad = int((oi != want_l).any(axis=1).sum())
        assert np.allclose(od, want_d, rtol=1e-5, atol=1e-6)
        report["oracle_parity"] = f"{a.batch} queries, {bad} differ only by distance ties within 1e-5"
print(json.dumps(report))
