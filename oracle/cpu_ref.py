"""CPU oracle for the datanode exact-kNN search path and the coordinator merge.

TEST INFRASTRUCTURE ONLY.  Nothing under ``distributed-vector-database_b200/`` may import
this module; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs do, and there only as the checker / the timed CPU baseline.

PARITY UNPINNED.  The reference repository ships no tests, golden vectors or known-answer
fixtures for the distance arithmetic; the arithmetic lives in the
third-party ``hnswlib`` package (nmslib/hnswlib, C++ header-only + pybind11; **not vendored
and not version-pinned** by the reference -- its requirements.txt does not list it; the
restatement below follows the published v0.8.0 sources), which is not installable in this
environment.  What IS pinned: the WAL record schema and replay semantics, against the ten
records checked in under ``Static/wal/node_1/`` (see ``replay_wal_records``).

What each function restates (paths relative to the reference checkout):

* ``l2sqr`` / ``inner_product_distance`` / ``normalize_rows``
      hnswlib ``space_l2.h`` (``L2SqrSIMD16ExtAVX``), ``space_ip.h``
      (``InnerProductSIMD16ExtAVX`` -> ``1 - ip``) and ``python_bindings/bindings.cpp``
      (``normalize_vector``: ``x * 1/(sqrtf(sum x^2) + 1e-30f)``), reached from
      ``src/datanode/handler.py:46`` (space), ``:268`` (add) and ``:364`` (query).
      The 16-floats-per-iteration AVX variant keeps 8 fp32 partial sums, adds two 8-lane
      products per iteration and reduces the 8 lanes left to right at the end.
* ``knn_exact``            what ``hnswlib.BFIndex.knn_query`` returns: every live row scored,
      ascending ``(distance, label)`` (``bruteforce.h``: ``std::pair`` ordering in the
      result heap => ties go to the smaller label).
* ``datanode_search``      ``VectorNodeHandler.search`` ``src/datanode/handler.py:344-408``.
* ``coordinator_merge``    ``CoordinatorHandler.search`` ``src/coordinator/handler.py:180-225``.
* ``get_shard_id`` / ``assign_shards_to_nodes``  ``src/utils/shared_utils.py:4-21``.
* ``replay_wal_records``   ``WALManager.replay`` ``src/utils/wal_manager.py:116-182``.
* ``synth_rows``           not in the reference: the counter-based synthetic generator of
      SURVEY.md §8d, bit-identical to the CUDA fill kernel (integer arithmetic + correctly
      rounded double sqrt/divide), so any row of a 10M-row device-generated set can be
      regenerated here.
"""
from __future__ import annotations

import hashlib
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

METRICS = ("l2", "ip", "cosine")

# --------------------------------------------------------------------------------------
# synthetic data (shared definition with csrc/synth.cuh)
# --------------------------------------------------------------------------------------
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)
SEED_DB = 0xD5B200
SEED_QUERY = 0xC0FFEE


def _splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def synth_rows(seed: int, row_start: int, n: int, dim: int) -> np.ndarray:
    """Rows ``row_start .. row_start+n`` of the synthetic unit-norm set ``seed``.

    element(row, col): h = splitmix64(splitmix64(seed*GOLDEN + row) + col);
    v = sum of the four 16-bit fields of h - 131070 (Irwin-Hall(4), zero mean, integer);
    row = float32( v / sqrt(sum v^2) ) with the sum exact in int64 and sqrt/divide in fp64.
    """
    rows = np.arange(row_start, row_start + n, dtype=np.uint64)
    with np.errstate(over="ignore"):
        base = _splitmix64((np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15) + rows) & _M64)
        h = _splitmix64((base[:, None] + np.arange(dim, dtype=np.uint64)[None, :]) & _M64)
    f = np.uint64(0xFFFF)
    v = ((h & f) + ((h >> np.uint64(16)) & f) + ((h >> np.uint64(32)) & f)
         + (h >> np.uint64(48))).astype(np.int64) - 131070
    ss = (v * v).sum(axis=1)
    out = v.astype(np.float64) / np.sqrt(ss.astype(np.float64))[:, None]
    return out.astype(np.float32)


# --------------------------------------------------------------------------------------
# hnswlib distance definitions (fp32, AVX summation order)
# --------------------------------------------------------------------------------------
def _lanes8_sum(prod: np.ndarray) -> np.ndarray:
    """Sum ``prod[..., D]`` the way the AVX kernels do: 8 running fp32 lanes, then lane 0..7
    added left to right.  A tail (D % 16 != 0) is added sequentially afterwards, as
    ``L2SqrSIMD16ExtResiduals`` does."""
    d = prod.shape[-1]
    d16 = (d // 16) * 16
    res = np.zeros(prod.shape[:-1], dtype=np.float32)
    if d16:
        lanes = np.zeros(prod.shape[:-1] + (8,), dtype=np.float32)
        for j in range(0, d16, 8):
            lanes = lanes + prod[..., j:j + 8]
        res = lanes[..., 0]
        for i in range(1, 8):
            res = res + lanes[..., i]
    if d16 < d:
        tail = np.zeros(prod.shape[:-1], dtype=np.float32)
        for j in range(d16, d):
            tail = tail + prod[..., j]
        res = res + tail
    return res.astype(np.float32)


def l2sqr(q: np.ndarray, rows: np.ndarray) -> np.ndarray:
    """Squared L2 (no sqrt), space ``'l2'``.  q [D] fp32, rows [N, D] fp32 -> [N] fp32."""
    q = np.asarray(q, dtype=np.float32)
    rows = np.asarray(rows, dtype=np.float32)
    diff = rows - q[None, :]
    return _lanes8_sum(diff * diff)


def inner_product_distance(q: np.ndarray, rows: np.ndarray) -> np.ndarray:
    """``1 - sum q_i d_i``, space ``'ip'`` (and ``'cosine'`` after normalisation)."""
    q = np.asarray(q, dtype=np.float32)
    rows = np.asarray(rows, dtype=np.float32)
    return (np.float32(1.0) - _lanes8_sum(rows * q[None, :])).astype(np.float32)


def normalize_rows(x: np.ndarray) -> np.ndarray:
    """hnswlib ``normalize_vector``: sequential fp32 sum of squares, then
    ``x * (1 / (sqrtf(norm) + 1e-30f))``."""
    x = np.atleast_2d(np.asarray(x, dtype=np.float32))
    norm = np.zeros(x.shape[0], dtype=np.float32)
    for j in range(x.shape[1]):
        norm = norm + x[:, j] * x[:, j]
    inv = np.float32(1.0) / (np.sqrt(norm, dtype=np.float32) + np.float32(1e-30))
    return (x * inv[:, None]).astype(np.float32)


def prepare_rows(rows: np.ndarray, metric: str, store_dtype: str = "f32") -> np.ndarray:
    """What the index holds after ``add_items``: normalised for cosine; optionally rounded to
    fp16 storage (a build extension -- the reference stores fp32 only, handler.py:224)."""
    rows = np.atleast_2d(np.asarray(rows, dtype=np.float32))
    if metric == "cosine":
        rows = normalize_rows(rows)
    if store_dtype == "f16":
        rows = rows.astype(np.float16).astype(np.float32)
    return rows


def distances(q: np.ndarray, stored: np.ndarray, metric: str) -> np.ndarray:
    """fp32 distances of one query against already-prepared rows."""
    q = np.asarray(q, dtype=np.float32).reshape(-1)
    if metric == "l2":
        return l2sqr(q, stored)
    if metric == "ip":
        return inner_product_distance(q, stored)
    if metric == "cosine":
        return inner_product_distance(normalize_rows(q)[0], stored)
    raise ValueError(f"unknown metric {metric!r}")


def distances_f64(q: np.ndarray, stored: np.ndarray, metric: str) -> np.ndarray:
    """Same definitions in float64 -- the tie/near-tie arbiter for tolerance checks."""
    q = np.asarray(q, dtype=np.float32).reshape(-1)
    if metric == "cosine":
        q = normalize_rows(q)[0]
    q64 = q.astype(np.float64)
    s64 = np.asarray(stored, dtype=np.float32).astype(np.float64)
    if metric == "l2":
        d = s64 - q64[None, :]
        return (d * d).sum(axis=1)
    return 1.0 - s64 @ q64


def topk_by_dist_label(dist: np.ndarray, labels: np.ndarray, k: int) -> Tuple[np.ndarray, np.ndarray]:
    """Ascending (distance, label); the first k."""
    order = np.lexsort((labels, dist))[:k]
    return labels[order], dist[order]


def knn_exact(queries: np.ndarray, stored: np.ndarray, labels: np.ndarray, k: int, metric: str,
              deleted: Optional[Iterable[int]] = None) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Exact k-NN over all live rows.  Returns (labels int64 [nq,k], dist fp32 [nq,k],
    counts int32 [nq]); short rows are padded with label -1 / distance +inf."""
    queries = np.atleast_2d(np.asarray(queries, dtype=np.float32))
    labels = np.asarray(labels, dtype=np.int64)
    live = np.ones(len(labels), dtype=bool)
    if deleted is not None:
        dl = np.fromiter(deleted, dtype=np.int64)
        if dl.size:
            live &= ~np.isin(labels, dl)
    s = stored[live]
    lab = labels[live]
    nq = queries.shape[0]
    out_l = np.full((nq, k), -1, dtype=np.int64)
    out_d = np.full((nq, k), np.inf, dtype=np.float32)
    cnt = np.zeros(nq, dtype=np.int32)
    for i in range(nq):
        if len(lab) == 0:
            continue
        d = distances(queries[i], s, metric)
        ll, dd = topk_by_dist_label(d, lab, k)
        out_l[i, :len(ll)] = ll
        out_d[i, :len(dd)] = dd
        cnt[i] = len(ll)
    return out_l, out_d, cnt


# --------------------------------------------------------------------------------------
# fast path for large N (used by bench cpu_baseline and big parity checks): BLAS scores,
# exact re-score of a generous candidate set with the definitions above.
# --------------------------------------------------------------------------------------
def knn_exact_fast(queries: np.ndarray, stored: np.ndarray, labels: np.ndarray, k: int, metric: str,
                   slack: int = 64) -> Tuple[np.ndarray, np.ndarray]:
    """Same result as ``knn_exact`` (no deletions) but O(N) BLAS first: take the k+slack best
    by a float32 GEMM score, then re-score those with the hnswlib-order arithmetic.  The
    GEMM score differs from the definition by ~1e-6, far less than the k..k+slack gap on the
    synthetic sets; callers that need a proof use ``knn_exact``."""
    queries = np.atleast_2d(np.asarray(queries, dtype=np.float32))
    labels = np.asarray(labels, dtype=np.int64)
    qn = normalize_rows(queries) if metric == "cosine" else queries
    out_l = np.empty((queries.shape[0], k), dtype=np.int64)
    out_d = np.empty((queries.shape[0], k), dtype=np.float32)
    sq = (stored * stored).sum(axis=1) if metric == "l2" else None
    kk = min(k + slack, stored.shape[0])
    for i0 in range(0, queries.shape[0], 256):
        qb = qn[i0:i0 + 256]
        ip = qb @ stored.T
        score = (sq[None, :] - 2.0 * ip) if metric == "l2" else -ip
        cand = np.argpartition(score, kk - 1, axis=1)[:, :kk]
        for j in range(qb.shape[0]):
            c = cand[j]
            d = distances(queries[i0 + j], stored[c], metric)
            ll, dd = topk_by_dist_label(d, labels[c], k)
            out_l[i0 + j], out_d[i0 + j] = ll, dd
    return out_l, out_d


# --------------------------------------------------------------------------------------
# datanode + coordinator semantics
# --------------------------------------------------------------------------------------
class DatanodeModel:
    """State of one ``VectorNodeHandler`` as far as search results depend on it
    (src/datanode/handler.py:222-342): monotonically increasing ids, overwrite = tombstone
    the old id + append a new one, delete = tombstone."""

    def __init__(self, dim: int = 512, metric: str = "l2", store_dtype: str = "f32"):
        self.dim, self.metric, self.store_dtype = dim, metric, store_dtype
        self.rows: List[np.ndarray] = []
        self.keys: List[str] = []          # id -> key
        self.meta: List[dict] = []
        self.raw: List[np.ndarray] = []
        self.key_to_id: Dict[str, int] = {}
        self.deleted: set = set()

    def put(self, key: str, vector: Sequence[float], metadata: Optional[dict] = None) -> bool:
        vec = np.array(vector, dtype=np.float32)              # handler.py:224
        if vec.ndim != 1 or vec.shape[0] != self.dim:         # :228-232
            return False
        old = self.key_to_id.get(key, -1)                     # :254-261
        if old != -1:
            self.deleted.add(old)
        new_id = len(self.rows)                               # :264
        self.rows.append(prepare_rows(vec[None, :], self.metric, self.store_dtype)[0])
        self.raw.append(vec)
        self.keys.append(key)
        self.meta.append(dict(metadata or {}))
        self.key_to_id[key] = new_id
        return True

    def delete(self, key: str) -> bool:                       # :323-342
        hid = self.key_to_id.get(key, -1)
        if hid == -1:
            return False
        self.deleted.add(hid)
        del self.key_to_id[key]
        return True

    def count(self) -> int:
        return len(self.rows)


def datanode_search(node: DatanodeModel, query_vector: Sequence[float], top_k: int,
                    reference_quirks: bool = False):
    """``VectorNodeHandler.search`` (handler.py:344-408) over an exact index.

    Returns (success, keys, scores).  ``top_k <= 0`` -> 5 (:346); empty index -> success with
    empty lists (:353-354); ``k = min(top_k, count)`` (:357); ask the index for ``2k`` (:364),
    drop tombstoned ids (:378) and stop at ``top_k`` (:402).  With ``reference_quirks`` the
    ``2k > count`` case fails the way hnswlib's RuntimeError makes the reference fail
    (:366-369); without it the exact index simply returns what is live (deliberate deviation,
    SURVEY.md §8b)."""
    q = np.array(query_vector, dtype=np.float32).reshape(1, -1)
    top_k = top_k if top_k > 0 else 5
    count = node.count()
    if count == 0:
        return True, [], []
    k = min(top_k, count)
    want = 2 * k
    if want > count:
        if reference_quirks:
            return False, [], []
        want = count
    stored = np.stack(node.rows)
    labels = np.arange(count, dtype=np.int64)
    # hnswlib returns the 2k nearest INCLUDING ids the handler has tombstoned (the reference
    # never calls mark_deleted); the handler filters afterwards.
    ll, dd, cnt = knn_exact(q, stored, labels, want, node.metric)
    keys, scores = [], []
    for i in range(int(cnt[0])):
        hid = int(ll[0, i])
        if hid in node.deleted:
            continue
        keys.append(node.keys[hid])
        scores.append(float(dd[0, i]))
        if len(keys) >= top_k:
            break
    return True, keys, scores


def datanode_search_exact_live(node: DatanodeModel, query_vector: Sequence[float], top_k: int):
    """The north-star target: top_k over LIVE rows only (tombstones masked inside the scan, so
    results do not run short when more than k of the 2k nearest are deleted)."""
    q = np.array(query_vector, dtype=np.float32).reshape(1, -1)
    top_k = top_k if top_k > 0 else 5
    if node.count() == 0:
        return True, [], []
    stored = np.stack(node.rows)
    labels = np.arange(node.count(), dtype=np.int64)
    ll, dd, cnt = knn_exact(q, stored, labels, top_k, node.metric, deleted=node.deleted)
    n = int(cnt[0])
    return True, [node.keys[int(i)] for i in ll[0, :n]], [float(x) for x in dd[0, :n]]


def coordinator_merge(per_node: Sequence[Tuple[Sequence[str], Sequence[float]]], top_k: int):
    """``CoordinatorHandler.search`` merge (coordinator/handler.py:200-216): concatenate in
    node order, first occurrence of a key wins, stable ascending sort by score, slice."""
    all_keys: List[str] = []
    all_scores: List[float] = []
    seen = set()
    for keys, scores in per_node:
        for k_, s_ in zip(keys, scores):
            if k_ in seen:
                continue
            seen.add(k_)
            all_keys.append(k_)
            all_scores.append(s_)
    if not all_scores:
        return [], []
    order = sorted(range(len(all_scores)), key=lambda i: all_scores[i])[:top_k]
    return [all_keys[i] for i in order], [all_scores[i] for i in order]


def merge_topk_by_id(dist: np.ndarray, ids: np.ndarray, k: int) -> Tuple[np.ndarray, np.ndarray]:
    """The cross-GPU form of the merge: ``dist``/``ids`` [G, nq, kk] (id -1 = padding) ->
    ascending (distance, id) top-k per query.  Same as ``coordinator_merge`` when scores do
    not tie across nodes; ties go to the smaller id (north_star) instead of node order."""
    g, nq, kk = dist.shape
    out_d = np.full((nq, k), np.inf, dtype=np.float32)
    out_i = np.full((nq, k), -1, dtype=np.int64)
    for q in range(nq):
        d = dist[:, q, :].reshape(-1)
        i = ids[:, q, :].reshape(-1)
        m = i >= 0
        ll, dd = topk_by_dist_label(d[m], i[m], k)
        out_i[q, :len(ll)], out_d[q, :len(dd)] = ll, dd
    return out_d, out_i


def get_shard_id(key: str, shard_count: int) -> int:
    """src/utils/shared_utils.py:4-7."""
    return int(hashlib.md5(key.encode()).hexdigest(), 16) % shard_count


def assign_shards_to_nodes(nodes: list, shard_count: int, replica_count: int = 2) -> dict:
    """src/utils/shared_utils.py:9-21."""
    mapping = {}
    if not nodes:
        return mapping
    for shard_id in range(shard_count):
        mapping[shard_id] = {
            "master": nodes[shard_id % len(nodes)],
            "slaves": [nodes[(shard_id + i) % len(nodes)] for i in range(1, replica_count + 1)],
        }
    return mapping


def replay_wal_records(records: Iterable[dict], after_ts: int = 0) -> List[dict]:
    """``WALManager.replay`` / ``replay_incremental`` reduction (wal_manager.py:131-175,
    :200-240): keep the LAST op per key, emitted in FIRST-appearance order of the key
    (python dict semantics); records with ``timestamp <= after_ts`` are skipped."""
    unique: Dict[str, dict] = {}
    for rec in records:
        if after_ts and rec["timestamp"] <= after_ts:
            continue
        unique[rec["key"]] = rec
    return list(unique.values())


# --------------------------------------------------------------------------------------
# tolerance-aware comparison used by every parity test
# --------------------------------------------------------------------------------------
def check_topk(got_ids: np.ndarray, got_dist: np.ndarray, q: np.ndarray, stored: np.ndarray,
               labels: np.ndarray, k: int, metric: str, deleted: Optional[Iterable[int]] = None,
               rtol: float = 1e-5) -> Optional[str]:
    """None if ``got`` is an acceptable exact top-k for query ``q``: ids identical to the
    oracle's, except that rows whose float64 distances lie within ``rtol*max(1,|d|)`` of each
    other may swap (a distance tie within tolerance); every reported distance within
    ``rtol*max(1,|d|)`` of the float64 definition.  Otherwise a message."""
    labels = np.asarray(labels, dtype=np.int64)
    live = np.ones(len(labels), dtype=bool)
    if deleted is not None:
        dl = np.fromiter(deleted, dtype=np.int64)
        if dl.size:
            live &= ~np.isin(labels, dl)
    d64 = distances_f64(q, stored, metric)
    d64_live = np.where(live, d64, np.inf)
    order = np.lexsort((labels, d64_live))
    n_live = int(live.sum())
    kk = min(k, n_live)
    got_ids = np.asarray(got_ids)[:kk]
    got_dist = np.asarray(got_dist, dtype=np.float64)[:kk]
    if len(got_ids) < kk:
        return f"short result: {len(got_ids)} < {kk}"
    if len(set(got_ids.tolist())) != kk:
        return "duplicate ids in result"
    pos = {int(l): i for i, l in enumerate(labels)}
    for r in range(kk):
        gid = int(got_ids[r])
        if gid not in pos or not live[pos[gid]]:
            return f"rank {r}: id {gid} is not a live row"
        true_d = d64[pos[gid]]
        tol = rtol * max(1.0, abs(true_d))
        if abs(got_dist[r] - true_d) > tol:
            return f"rank {r}: id {gid} distance {got_dist[r]!r} vs definition {true_d!r}"
        want_d = d64_live[order[r]]
        if gid != int(labels[order[r]]) and abs(true_d - want_d) > 2 * tol:
            return (f"rank {r}: id {gid} (d={true_d!r}) where oracle has {int(labels[order[r]])} "
                    f"(d={want_d!r}) -- not a tie within tolerance")
    if kk and np.any(np.diff(got_dist) < 0):
        return "distances not ascending"
    return None
