#!/usr/bin/env python
"""One shard of a BASELINE config on one GPU: build it from the on-device synthetic generator, then check the
size-independent properties the parity tests use at full size and time both search paths.

  python tools/config_check.py --rows 1250000 --dim 512 --store f32 --metric l2 --k 100 --batch 4096   (config 3 shard)
  python tools/config_check.py --rows 12500000 --dim 768 --store f16 --metric ip --k 10 --batch 1024    (config 4 shard)

Checks: (1) stored rows used as queries find themselves first (distance ~0 for l2/cosine); (2) results are
sorted and full; (3) the batched tensor-core path and the single-query scan path (different kernels, same
summation order in the re-rank) return bit-identical ids and distances for the same queries; (4) a sample of
queries equals the CPU oracle on the rows regenerated on the host (only when rows <= --oracle-rows)."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import dvdb_b200 as vdb
from oracle import cpu_ref as R

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, required=True)
ap.add_argument("--dim", type=int, default=512)
ap.add_argument("--store", default="f32")
ap.add_argument("--metric", default="cosine")
ap.add_argument("--k", type=int, default=10)
ap.add_argument("--batch", type=int, default=1024)
ap.add_argument("--oracle-rows", type=int, default=2_000_000)
ap.add_argument("--steps", type=int, default=5)
a = ap.parse_args()

t0 = time.time()
ix = vdb.Index(a.metric, a.dim, store_dtype=a.store)
ix.init_index(a.rows)
ix.add_synthetic(R.SEED_DB, 0, a.rows)
torch.cuda.synchronize()
build_s = time.time() - t0
elem = 2 if a.store == "f16" else 4
shard_gb = a.rows * ix.get_stat("ld") * elem / 1e9
out = {"config": vars(a), "shard_gb": shard_gb, "shadow_plane": bool(ix.get_stat("shadow")), "build_s": build_s}

probe = [0, a.rows // 3, a.rows - 1]
q = np.concatenate([R.synth_rows(R.SEED_DB, r, 1, a.dim) for r in probe] + [R.synth_rows(R.SEED_QUERY, 0, a.batch - len(probe), a.dim)])
lab, dist, cnt = ix.knn_query_padded(q, a.k)
assert (cnt == a.k).all() and (np.diff(dist, axis=1) >= 0).all(), "results must be full and sorted"
if a.metric != "ip":
    for i, r in enumerate(probe):
        assert lab[i, 0] == r and abs(dist[i, 0]) < 1e-5, f"row {r} does not find itself: {lab[i,:3]} {dist[i,:3]}"
out["self_query"] = "ok"
assert ix.get_stat("tensor_batches") >= 1, "the batch did not take the tensor path"
ns = 24
for i in range(ns):
    l1, d1, _ = ix.knn_query_padded(q[i:i + 1], a.k)
    assert np.array_equal(l1[0], lab[i]) and np.array_equal(d1[0], dist[i]), f"query {i}: tensor path != scan path"
out["tensor_equals_scan_bitwise"] = f"{ns} queries"
out["fallback_queries"] = ix.get_stat("fallback_queries")
if a.rows <= a.oracle_rows:
    from oracle import c_ref
    rows = c_ref.synth_rows(R.SEED_DB, 0, a.rows, a.dim)
    stored = c_ref.normalize(rows) if a.metric == "cosine" else rows
    if a.store == "f16":
        stored = stored.astype(np.float16).astype(np.float32)
    m = 64
    want_l, want_d, _ = c_ref.knn(q[:m], stored, None, a.k, a.metric)
    for i in range(m):
        if not np.array_equal(lab[i], want_l[i]):
            u = np.unique(np.concatenate([want_l[i], lab[i]]))
            msg = R.check_topk(lab[i], dist[i], q[i], stored[u], u, a.k, a.metric, rtol=1e-5)
            assert msg is None, f"query {i}: {msg}"
    out["oracle_parity"] = f"{m} queries, ids identical up to 1e-5 distance ties"

# timings (host-buffer API, pinned)
def timed(fn, n):
    fn(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / n
qp = vdb.pinned_empty(q.shape, np.float32); qp[:] = q
t_batch = timed(lambda: ix.knn_query_padded(qp, a.k), a.steps)
t_one = timed(lambda: ix.knn_query_padded(qp[:1], a.k), 10 * a.steps)
out["batch_ms"] = 1e3 * t_batch
out["batch_qps"] = a.batch / t_batch
out["batch_tflops"] = 2.0 * a.batch * a.rows * a.dim / t_batch / 1e12
out["single_query_ms"] = 1e3 * t_one
out["single_query_gbs"] = shard_gb / t_one
print(json.dumps(out))
