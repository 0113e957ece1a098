#!/usr/bin/env python
"""Datanode-handler throughput under concurrent single-query requests (the reference's serving shape: a pool of
Thrift worker threads, one query per SearchRequest): one index query per request under the handler lock, as in
the reference, against the micro-batched handler (concurrent requests coalesced into tensor-path batches).

  python tools/serve_threads.py [--rows 1000000] [--threads 1 5 16 64]"""
import argparse, json, os, sys, threading, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import dvdb_b200 as vdb
from oracle import cpu_ref as R

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=1_000_000)
ap.add_argument("--dim", type=int, default=512)
ap.add_argument("--threads", type=int, nargs="+", default=[1, 5, 16, 64])
ap.add_argument("--per-thread", type=int, default=40)
ap.add_argument("--wait-us", type=float, default=200.0)
a = ap.parse_args()

out = {"rows": a.rows, "dim": a.dim, "per_thread": a.per_thread, "wait_us": a.wait_us, "runs": []}
with tempfile.TemporaryDirectory() as root:
    for mode, wait in (("one query per request (reference behaviour)", None), ("micro-batched", a.wait_us * 1e-6)):
        h = vdb.GpuVectorNodeHandler("bench_" + ("mb" if wait is not None else "plain"), storage_root=root, space="cosine",
                                     dim=a.dim, max_elements=a.rows, checkpoint_every=0, fsync=False, micro_batch_wait_s=wait)
        # rows straight into the index (the key table is not what is measured: answers carry no keys here)
        h.hnsw_index.add_synthetic(R.SEED_DB, 0, a.rows)
        qs = R.synth_rows(R.SEED_QUERY, 0, 256, a.dim)
        reqs = [vdb.SearchRequest(query_vector=q.tolist(), top_k=10) for q in qs]
        h.search(reqs[0])
        for nt in a.threads:
            def worker(t):
                for j in range(a.per_thread):
                    h.search(reqs[(t * a.per_thread + j) % len(reqs)])
            ts = [threading.Thread(target=worker, args=(t,)) for t in range(nt)]
            t0 = time.perf_counter()
            [t.start() for t in ts]
            [t.join() for t in ts]
            dt = time.perf_counter() - t0
            run = {"mode": mode, "threads": nt, "requests": nt * a.per_thread, "qps": nt * a.per_thread / dt,
                   "ms_per_request_per_thread": 1e3 * dt / a.per_thread}
            if h._batcher is not None:
                run["requests_per_index_query"] = h._batcher.requests / max(h._batcher.batches, 1)
                h._batcher.requests = h._batcher.batches = 0
            out["runs"].append(run)
            print(json.dumps(run), flush=True)
        h.close()
print(json.dumps(out))
