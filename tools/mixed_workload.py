#!/usr/bin/env python
"""BASELINE config 5 (mixed insert + search through the WAL on 4 shards, cosine top-10, batch 256), scaled.

Four `GpuVectorNodeHandler`s (one shard each, md5 routing through `LocalCoordinator`'s rule) on one GPU.  The
database grows from --start to --end rows in --insert-batch puts (`put_batch`: one add_items + one WAL group
commit per shard); after every growth step a batch of 256 queries is searched on every shard
(`search_batch`, the tensor path) and the per-shard lists are merged with the GPU merge kernel.  Reports search
queries/s against N, insert rows/s and WAL bytes/s, then checks the final state against the CPU oracle and a
cold restart (checkpoint + WAL replay) against the live state.

The reference's record keeps every vector as a JSON list in the WAL and as Python objects in the key table
(~16 KB per row in CPython), so the full 1M -> 5M run is a memory exercise for the host; the default here is
1/20 scale.  `--level index` runs the same growth/search schedule on bare `Index` shards (no per-key Python
objects, no WAL) at any size."""
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
    root = tempfile.mkdtemp(prefix="vdb_mixed_")
    try:
        handlers = {f"node_{s}": vdb.GpuVectorNodeHandler(f"node_{s}", root, space="cosine", dim=a.dim,
                                                          max_elements=a.end, checkpoint_every=0, fsync=False)
                    for s in range(a.shards)}
        node_ids = list(handlers)
        mapping = vdb.assign_shards_to_nodes(node_ids, a.shards)
        # global row id <-> (shard, local hnsw id): results are compared through keys
        def put_rows(lo, hi):
            rows = c_ref.synth_rows(R.SEED_DB, lo, hi - lo, a.dim)
            per = {n: [] for n in node_ids}
            for r, key in enumerate(keys_of(lo, hi)):
                per[mapping[vdb.get_shard_id(key, a.shards)]["master"]].append(
                    vdb.VectorData(key=key, vector=rows[r].tolist(), metadata={"row": lo + r}))
            t = time.perf_counter()
            for n, items in per.items():
                resp = handlers[n].put_batch(items)
                assert resp.success, resp.message
            return time.perf_counter() - t

        def search_all():
            outs = [handlers[n].search_batch(queries, a.k) for n in node_ids]
            merged = []
            for qi in range(a.batch):
                cand = [(s, k) for ks, ss in outs for k, s in zip(ks[qi], ss[qi])]
                cand.sort()
                merged.append(cand[:a.k])
            return merged

        n = 0
        wal0 = 0
        while n < a.end:
            hi = min(a.end, n + (a.start if n == 0 else a.insert_batch))
            dt_ins = put_rows(n, hi)
            wal1 = sum(dir_bytes(handlers[x].wal_dir) for x in node_ids)
            search_all()
            t = time.perf_counter()
            for _ in range(a.searches_per_step):
                res = search_all()
            dt_s = (time.perf_counter() - t) / a.searches_per_step
            report["steps"].append({"rows": hi, "insert_rows_per_s": (hi - n) / dt_ins, "wal_MB_per_s": (wal1 - wal0) / dt_ins / 1e6,
                                    "search_qps": a.batch / dt_s, "search_ms": 1e3 * dt_s})
            n, wal0 = hi, wal1
        # final state against the oracle
        stored = c_ref.normalize(c_ref.synth_rows(R.SEED_DB, 0, a.end, a.dim))
        want_l, want_d, _ = c_ref.knn(queries, stored, None, a.k, "cosine")
        bad = 0
        for qi in range(a.batch):
            got = [int(k[1:]) for _, k in res[qi]]
            if got != want_l[qi].tolist():
                bad += 1
                gd = np.array([s for s, _ in res[qi]], np.float32)
                assert np.allclose(gd, want_d[qi], rtol=1e-5, atol=1e-6), f"query {qi}: {got} vs {want_l[qi]}"
        report["oracle_parity"] = f"{a.batch} queries, {bad} differ only by distance ties within 1e-5"
        # cold restart: checkpoint half-way is not used here (checkpoint_every=0) -> full WAL replay
        h0 = handlers[node_ids[0]]
        live_keys = sorted(h0._by_key)
        t = time.perf_counter()
        h0.hnsw_index.close()
        h0b = vdb.GpuVectorNodeHandler(node_ids[0], root, space="cosine", dim=a.dim, max_elements=a.end,
                                       checkpoint_every=0, fsync=False)
        report["wal_replay_s"] = time.perf_counter() - t
        assert sorted(h0b._by_key) == live_keys, "WAL replay lost or invented keys"
        k1, s1 = h0b.search_batch(queries[:16], a.k)
        k0 = [[k for _, k in r] for r in res[:16]]
        report["wal_replay"] = f"{len(live_keys)} keys restored on shard 0"
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
