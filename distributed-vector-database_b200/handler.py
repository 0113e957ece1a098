"""GpuVectorNodeHandler -- the datanode handler of the reference with the GPU index behind it.

Same public methods, argument meaning and Response behaviour as `VectorNodeHandler`
(reference src/datanode/handler.py:222 put, :323 delete, :344 search, :411 get, :156/:181
checkpoints), so the generated Thrift `VectorNodeService.Processor` can serve it unchanged.
What changes underneath:

 * `hnswlib.Index`  -> `index.Index` (exact search on the B200, tombstones masked in the scan, so
   the handler asks for `top_k` results instead of `2*top_k` and never runs short)
 * LevelDB JSON values + the O(N) id->key scan (handler.py:145-153) -> `kvstore.KeyStore`: O(1) id<->key
   maps in memory, raw vectors and metadata in two append-only files under `leveldb_data/` (a checkpoint
   records a position in them instead of copying the store)
 * `index.bin` is the flat GPU shard snapshot (vdb_save) instead of hnswlib's private format
 * the index is NOT rewritten to disk on every put (handler.py:302-304); durability comes from the
   WAL (group commit) + periodic checkpoints, as in the reference's own recovery path.

Out of scope here: ZooKeeper registration (handler.py:39) and the Thrift server bootstrap.
"""
from __future__ import annotations

import json
import os
import shutil
import threading
import time
from typing import Callable, Dict, List, Optional

import numpy as np

from .sharding import VECTOR_DIM
from .kvstore import KeyStore
from .ttypes import Response, SearchRequest, SearchResult, VectorData
from .wal import WALManager


def _default_index_factory(space: str, dim: int, max_elements: int, store_dtype: str, device: int):
    from .index import Index            # CUDA path; raises if the library / a GPU is missing
    ix = Index(space, dim, store_dtype=store_dtype, device=device)
    ix.init_index(max_elements=max_elements, ef_construction=128, M=32)      # handler.py:86
    return ix


class _MicroBatcher:
    """Coalesces concurrent single-query searches into one batched index call.

    The reference serves `search` from a pool of 5 Thrift worker threads (datanode/server.py:25-28) that all
    queue on one lock (handler.py:23, :348): five concurrent requests cost five index queries back to back.  Here
    the first thread to arrive becomes the leader: it waits `max_wait_s` for the others to enqueue, runs ONE
    `run_batch(queries[n, dim], k_max)` (the tensor-core path for n > 4) and hands every follower its rows; the
    first k entries of an exact top-k_max are the exact top-k, so requests with different k share a batch.
    Leadership passes to a waiting follower after each batch; no dedicated thread."""

    PEER_WINDOW_S = 5e-3        # "recently": another thread submitted within this long

    class _Item:
        __slots__ = ("vec", "k", "out", "err", "done")

        def __init__(self, vec, k):
            self.vec, self.k, self.out, self.err, self.done = vec, k, None, None, False

    def __init__(self, run_batch: Callable, max_batch: int = 256, max_wait_s: float = 2e-4):
        self._run, self.max_batch, self.max_wait_s = run_batch, max_batch, max_wait_s
        self._cv = threading.Condition()
        self._queue: List["_MicroBatcher._Item"] = []
        self._leader = False
        self._last_batch = 1
        self._last_tid, self._last_switch = threading.get_ident(), -1e9
        self.batches = 0            # statistics: index calls / requests served
        self.requests = 0

    def submit(self, vec: np.ndarray, k: int):
        """-> (labels[<=k], distances[<=k], count) of this query; raises what `run_batch` raised."""
        item = self._Item(vec, k)
        tid = threading.get_ident()
        with self._cv:
            self._queue.append(item)
            if tid != self._last_tid:                 # requests are coming from more than one thread
                self._last_tid, self._last_switch = tid, time.perf_counter()
        while True:
            with self._cv:
                while self._leader and not item.done:
                    self._cv.wait()
                if item.done:
                    return self._result(item)
                self._leader = True                   # nobody is serving: this thread does, for one batch
            self._lead_one_batch()                    # FIFO: its own item is in it unless max_batch were ahead

    def _lead_one_batch(self) -> None:
        batch: list = []
        try:
            # let the other worker threads enqueue -- only when there are any: requests arriving from ONE thread
            # (a lone client) never wait, and a queue as long as the last batch needs no waiting either
            if (self.max_wait_s > 0 and time.perf_counter() - self._last_switch < self.PEER_WINDOW_S
                    and len(self._queue) < max(self._last_batch, 2)):
                time.sleep(self.max_wait_s)
            with self._cv:
                batch = self._queue[:self.max_batch]
                del self._queue[:len(batch)]
            try:
                k_max = max(it.k for it in batch)
                labels, dist, counts = self._run(np.stack([it.vec for it in batch]), k_max)
                for r, it in enumerate(batch):
                    it.out = (labels[r, :it.k], dist[r, :it.k], min(int(counts[r]), it.k))
            except BaseException as e:                # every request of the batch sees the failure
                for it in batch:
                    it.err = e
        finally:
            with self._cv:
                for it in batch:
                    it.done = True
                self.batches += 1
                self.requests += len(batch)
                self._last_batch = len(batch)
                self._leader = False
                self._cv.notify_all()

    @staticmethod
    def _result(item):
        if item.err is not None:
            raise item.err
        return item.out


class GpuVectorNodeHandler:
    def __init__(self, node_id: str, storage_root: str = "./Static/local_storage", *, space: str = "l2",
                 dim: int = VECTOR_DIM, max_elements: int = 1_000_000, store_dtype: str = "f32", device: int = 0,
                 checkpoint_every: int = 2000, reference_quirks: bool = False, fsync: bool = True,
                 index_factory: Optional[Callable] = None, micro_batch_wait_s: Optional[float] = None,
                 micro_batch_max: int = 256, keep_checkpoints: int = 2):
        self.node_id = node_id
        self.index_lock = threading.RLock()                      # handler.py:23
        self.space, self.vector_dim, self.store_dtype, self.device = space, dim, store_dtype, device
        self.max_elements = max_elements
        self.checkpoint_every = checkpoint_every                 # handler.py:316
        self.keep_checkpoints = keep_checkpoints
        self.reference_quirks = reference_quirks
        self._factory = index_factory or _default_index_factory
        # directory layout of the reference (handler.py:27-36)
        self.local_storage_dir = os.path.join(storage_root, str(node_id))
        self.hnsw_index_dir = os.path.join(self.local_storage_dir, "hnsw_index")
        self.leveldb_dir = os.path.join(self.local_storage_dir, "leveldb_data")
        self.wal_dir = os.path.join(self.local_storage_dir, "wal")
        self.checkpoint_dir = os.path.join(self.local_storage_dir, "checkpoint")
        self.deleted_ids_path = os.path.join(self.local_storage_dir, "deleted_ids.json")
        for d in (self.hnsw_index_dir, self.leveldb_dir, self.wal_dir, self.checkpoint_dir):
            os.makedirs(d, exist_ok=True)
        self.wal_manager = WALManager(self.wal_dir, node_id=str(node_id), fsync=fsync)
        self.next_hnsw_id = 0
        self.deleted_ids: set = set()
        # key <-> id maps, raw vectors, metadata: replaces LevelDB (handler.py:288-297) and its O(N) reverse scan
        self.store = KeyStore(self.leveldb_dir, dim, fsync=fsync)
        self.hnsw_index = self._factory(space, dim, max_elements, store_dtype, device)
        # micro_batch_wait_s = None: every search is its own index query, as in the reference; a number (0 allowed):
        # concurrent searches are coalesced (see _MicroBatcher)
        self._batcher = (None if micro_batch_wait_s is None else
                         _MicroBatcher(self._locked_query, micro_batch_max, micro_batch_wait_s))
        self.load_from_checkpoint()

    # ---- key table ---------------------------------------------------------------------------
    def _get_hnsw_id_by_key(self, key: str) -> int:               # handler.py:136-143
        return self.store.id_of(key)

    def _get_key_by_hnsw_id(self, hnsw_id: int) -> str:           # handler.py:145-153, O(1) here
        return self.store.key_of(hnsw_id)

    def stored_vector(self, hnsw_id: int) -> np.ndarray:
        """The raw vector given at put time (WAL records of bulk inserts point here instead of carrying it)."""
        return np.array(self.store.vector(int(hnsw_id)), dtype=np.float32)

    # ---- checkpoints (handler.py:156-219) ------------------------------------------------------
    def save_checkpoint(self) -> str:
        """index.bin (flat shard snapshot) + leveldb_data/kv_pos.json (a POSITION in the append-only key store, not a
        copy of it) + deleted_ids.json + wal_pos.txt (timestamp, as the reference) + wal_seq.txt (the exact record
        number the checkpoint contains).  Older checkpoints beyond `keep_checkpoints` are removed."""
        with self.index_lock:
            ts = int(time.time() * 1000)
            path = os.path.join(self.checkpoint_dir, f"checkpoint_{ts}")
            while os.path.exists(path):
                ts += 1
                path = os.path.join(self.checkpoint_dir, f"checkpoint_{ts}")
            tmp = path + ".tmp"
            os.makedirs(tmp)
            if hasattr(self.hnsw_index, "save_image"):
                # the shard's append-only image under hnsw_index/ grows by the rows added since the last checkpoint;
                # index.bin is the small meta of this moment (count + tombstones): O(new rows), not O(shard)
                self.hnsw_index.save_image(self.hnsw_index_dir, os.path.join(tmp, "index.bin"))
            else:
                self.hnsw_index.save_index(os.path.join(tmp, "index.bin"))
            self.store.flush()
            kv_dir = os.path.join(tmp, "leveldb_data")
            os.makedirs(kv_dir, exist_ok=True)
            with open(os.path.join(kv_dir, "kv_pos.json"), "w", encoding="utf-8") as f:
                json.dump(dict(self.store.position(), next_hnsw_id=self.next_hnsw_id), f)
            with open(os.path.join(tmp, "deleted_ids.json"), "w", encoding="utf-8") as f:
                json.dump(sorted(self.deleted_ids), f)
            with open(os.path.join(tmp, "wal_pos.txt"), "w") as f:
                f.write(str(ts))
            with open(os.path.join(tmp, "wal_seq.txt"), "w") as f:
                f.write(str(self.wal_manager.last_seq))
            os.rename(tmp, path)                                  # a checkpoint directory is complete or absent
            if self.keep_checkpoints and self.keep_checkpoints > 0:
                done = sorted(d for d in os.listdir(self.checkpoint_dir) if d.startswith("checkpoint_") and not d.endswith(".tmp"))
                for old in done[:-self.keep_checkpoints]:
                    shutil.rmtree(os.path.join(self.checkpoint_dir, old), ignore_errors=True)
            return path

    def load_from_checkpoint(self) -> None:
        dirs = sorted(d for d in os.listdir(self.checkpoint_dir) if d.startswith("checkpoint_") and not d.endswith(".tmp"))
        if not dirs:
            self.store.rollback({"log_bytes": 0})                 # no snapshot: full replay (wal_manager.py:116)
            self.wal_manager.replay(self)
            return
        path = os.path.join(self.checkpoint_dir, dirs[-1])
        index_path = os.path.join(path, "index.bin")
        if os.path.exists(index_path):
            if hasattr(self.hnsw_index, "load_image") and self.hnsw_index.is_image_meta(index_path):
                self.hnsw_index.load_image(self.hnsw_index_dir, index_path, max_elements=self.max_elements)
            else:
                self.hnsw_index.load_index(index_path, max_elements=self.max_elements)     # handler.py:195
            self.next_hnsw_id = self.hnsw_index.get_current_count()
        pos_path = os.path.join(path, "leveldb_data", "kv_pos.json")
        if os.path.exists(pos_path):
            with open(pos_path, "r", encoding="utf-8") as f:
                self.store.rollback(json.load(f))                 # what came later is re-applied from the WAL
        del_path = os.path.join(path, "deleted_ids.json")
        if os.path.exists(del_path):
            with open(del_path, "r", encoding="utf-8") as f:
                self.deleted_ids = set(json.load(f))
        with open(os.path.join(path, "wal_pos.txt"), "r") as f:
            checkpoint_ts = int(f.read())
        after_seq = None
        seq_path = os.path.join(path, "wal_seq.txt")
        if os.path.exists(seq_path):
            with open(seq_path, "r") as f:
                after_seq = int(f.read())
        self.wal_manager.replay_incremental(self, checkpoint_ts, after_seq)  # handler.py:218

    # ---- put / delete (handler.py:222-342) -----------------------------------------------------
    def _ensure_room(self, n_more: int) -> None:
        """The reference rebuilds the graph without its tombstones when full (:240-251); a flat shard grows in
        place instead (no copy, searches keep running)."""
        need = self.hnsw_index.get_current_count() + n_more
        cap = self.hnsw_index.get_max_elements()
        if need > cap:
            self.hnsw_index.resize_index(max(cap * 2, need, 1024))

    def _log(self, fn, *args) -> None:
        """A WAL failure is logged, not raised: the operation has been applied (reference handler.py:306-311)."""
        try:
            fn(*args)
        except Exception as e:                                    # pragma: no cover - disk errors
            print(f"[{self.node_id}] WAL write failed: {e!r}")

    def put(self, data: VectorData, replay_mode: bool = False) -> Response:
        key = data.key
        vec = np.array(data.vector, dtype=np.float32)                       # :224
        metadata = data.metadata or {}
        if vec.ndim != 1 or vec.shape[0] != self.vector_dim:                # :228-232
            return Response(success=False, message=f"vector dim mismatch: expect {self.vector_dim}, got {vec.shape}")
        with self.index_lock:
            self._ensure_room(1)
            new_id = self.next_hnsw_id                                      # :264
            try:
                self.hnsw_index.add_items(vec.reshape(1, -1), np.array([new_id], dtype=np.int64))
            except RuntimeError as e:                                       # :272
                return Response(success=False, message=f"index add failed: {e}")
            old_id = self._get_hnsw_id_by_key(key)                          # :254-261: overwrite = tombstone + append
            if old_id != -1:                                                # (after the append: a failed add leaves
                self.deleted_ids.add(old_id)                                #  the old version in place)
                self.hnsw_index.mark_deleted([old_id])
            self.next_hnsw_id += 1                                          # :285
            self.store.put(new_id, key, vec, metadata)                      # :288-297
            if not replay_mode:
                self._log(self.wal_manager.write_log, "PUT", key,
                          data.vector if isinstance(data.vector, list) else vec.tolist(), metadata)
                if self.checkpoint_every and self.next_hnsw_id % self.checkpoint_every == 0:   # :316-317
                    self.save_checkpoint()
        return Response(success=True, message=f"key={key} 写入成功")

    def put_batch(self, items: List[VectorData]) -> Response:
        """Insert path for bulk loads (config 5): ONE add_items call, one key-store append and one WAL group commit
        (vectors logged by reference to the key store's raw-vector file, see wal.py)."""
        if not items:
            return Response(success=True, message="empty batch")
        return self.put_arrays([d.key for d in items], np.asarray([d.vector for d in items], dtype=np.float32),
                               [d.metadata for d in items])

    def put_arrays(self, keys: List[str], vecs: np.ndarray, metas: Optional[List[Optional[dict]]] = None) -> Response:
        """`put_batch` without the per-item objects: keys [n], float32 [n, dim], optional metadata dicts."""
        vecs = np.ascontiguousarray(vecs, dtype=np.float32)
        n = len(keys)
        if vecs.ndim != 2 or vecs.shape != (n, self.vector_dim):
            return Response(success=False, message=f"vector dim mismatch: expect {self.vector_dim}, got {vecs.shape}")
        if n == 0:
            return Response(success=True, message="empty batch")
        metas = metas if metas is not None else [None] * n
        with self.index_lock:
            self._ensure_room(n)
            first = self.next_hnsw_id
            ids = np.arange(first, first + n, dtype=np.int64)
            try:
                self.hnsw_index.add_items(vecs, ids)
            except RuntimeError as e:                                     # nothing else has been touched yet
                return Response(success=False, message=f"index add failed: {e}")
            # overwrite = tombstone + append; duplicates inside the batch: only the last occurrence stays live
            dead, last = [], {}
            for i, key in enumerate(keys):
                old = self._get_hnsw_id_by_key(key)
                if old != -1:
                    dead.append(old)
                if key in last:
                    dead.append(first + last[key])
                last[key] = i
            if dead:
                dead = sorted(set(dead))
                self.deleted_ids.update(dead)
                self.hnsw_index.mark_deleted(dead)
            self.next_hnsw_id += n
            if len(last) == n:
                self.store.put_batch(ids.tolist(), keys, vecs, metas)
            else:
                keep = sorted(last.values())
                self.store.put_batch([first + i for i in keep], [keys[i] for i in keep], vecs[keep], [metas[i] for i in keep])
            self._log(self.wal_manager.write_batch,
                      [("PUT", keys[i], None, metas[i] or {}, None, first + i) for i in range(n)])
            if self.checkpoint_every and first // self.checkpoint_every != self.next_hnsw_id // self.checkpoint_every:
                self.save_checkpoint()
        return Response(success=True, message=f"{n} keys written")

    def delete(self, key: str, replay_mode: bool = False) -> Response:
        with self.index_lock:
            hnsw_id = self._get_hnsw_id_by_key(key)                         # :326
            if hnsw_id == -1:
                return Response(success=False, message=f"key={key}不存在")
            self.deleted_ids.add(hnsw_id)                                   # :332
            self.hnsw_index.mark_deleted([hnsw_id])
            self.store.delete(key)
            if not replay_mode:
                self._log(self.wal_manager.write_log, "DELETE", key)        # :339
        return Response(success=True, message=f"key={key}删除成功")

    # ---- search / get (handler.py:344-428) -------------------------------------------------------
    def _locked_query(self, queries: np.ndarray, k: int):
        # The index is safe for concurrent searches next to one writer (searches snapshot the published row count),
        # so the query itself runs OUTSIDE the handler lock: a put never waits for a search and the other way round.
        return self.hnsw_index.knn_query_padded(queries, min(k, max(self.hnsw_index.get_current_count(), 1)))

    def search(self, req: SearchRequest) -> Response:
        query_vec = np.array(req.query_vector, dtype=np.float32).reshape(1, -1)     # :345
        top_k = req.top_k if req.top_k and req.top_k > 0 else 5                     # :346
        if self._batcher is not None:
            return self._search_coalesced(query_vec, top_k)
        current_count = self.hnsw_index.get_current_count()
        if current_count == 0:                                                       # :353-354
            return Response(success=True, search_result=SearchResult(keys=[], scores=[], vectors=[]))
        k = min(top_k, current_count)                                                # :357
        if self.reference_quirks and 2 * k > current_count:
            # hnswlib cannot fill k*2 results -> RuntimeError -> the reference answers this (:364-369)
            return Response(success=False, message="HNSW index corrupted, search aborted")
        try:
            labels, distances, counts = self.hnsw_index.knn_query_padded(query_vec, k)
        except RuntimeError:                                                         # :366
            return Response(success=False, message="HNSW index corrupted, search aborted")
        with self.index_lock:
            return self._search_response(labels[0], distances[0], int(counts[0]), top_k)

    def _search_response(self, labels, distances, count: int, top_k: int) -> Response:
        """labels -> keys, vectors, metadata (handler.py:375-408); caller holds index_lock"""
        keys, vectors, scores = [], [], []
        for i in range(count):                                                       # :375
            hnsw_id = int(labels[i])
            if hnsw_id < 0 or hnsw_id in self.deleted_ids:                           # :378 (already masked on the GPU)
                continue
            key = self._get_key_by_hnsw_id(hnsw_id)                                  # :382
            if not key:
                continue
            keys.append(key)
            vectors.append(VectorData(key=key, vector=self.store.vector(hnsw_id).tolist(),
                                      metadata=self.store.metadata(hnsw_id)))        # :387-398
            scores.append(float(distances[i]))                                       # :393
            if len(keys) >= top_k:                                                   # :402
                break
        return Response(success=True, search_result=SearchResult(keys=keys, scores=scores, vectors=vectors))

    def _search_coalesced(self, query_vec: np.ndarray, top_k: int) -> Response:
        """`search` with concurrent requests sharing one index query.  The handler lock is taken only for the key
        lookups of each request; a put/delete that lands between the query and the lookups can only remove a label
        from this request's answer (the lookups skip it)."""
        current_count = self.hnsw_index.get_current_count()
        if current_count == 0:
            return Response(success=True, search_result=SearchResult(keys=[], scores=[], vectors=[]))
        k = min(top_k, current_count)
        if self.reference_quirks and 2 * k > current_count:
            return Response(success=False, message="HNSW index corrupted, search aborted")
        try:
            labels, distances, count = self._batcher.submit(query_vec[0], k)
        except RuntimeError:
            return Response(success=False, message="HNSW index corrupted, search aborted")
        with self.index_lock:
            return self._search_response(labels, distances, count, top_k)

    def search_ids(self, queries, top_k: int):
        """Array form of `search` for co-located coordinators (coordinator.LocalCoordinator.search_batch): queries
        float32 [nq, dim] -> (labels int64 [nq, k] (-1 padded), distances float32 [nq, k] (+inf padded)) with
        k = top_k.  Tombstones are masked on the GPU; ids map to keys with `keys_of`."""
        q = np.ascontiguousarray(np.asarray(queries, dtype=np.float32))
        k = int(top_k) if top_k and top_k > 0 else 5
        nq = len(q)
        count = self.hnsw_index.get_current_count()
        if count == 0 or nq == 0:
            return np.full((nq, k), -1, dtype=np.int64), np.full((nq, k), np.inf, dtype=np.float32)
        kk = min(k, count)
        labels, distances, _ = self.hnsw_index.knn_query_padded(q, kk)
        if kk < k:
            labels = np.concatenate([labels, np.full((nq, k - kk), -1, dtype=np.int64)], axis=1)
            distances = np.concatenate([distances, np.full((nq, k - kk), np.inf, dtype=np.float32)], axis=1)
        return labels, distances

    def keys_of(self, labels) -> List[str]:
        """ids -> keys ('' for padding / ids deleted since the query)."""
        with self.index_lock:
            return [self.store.key_of(int(h)) if h >= 0 and int(h) not in self.deleted_ids else "" for h in labels]

    def search_batch(self, queries, top_k: int):
        """Additive (the IDL has one query per SearchRequest, vector_db.thrift:23-28): many queries in one
        call so the tensor-core path is used.  Returns (keys[nq][<=k], scores[nq][<=k])."""
        q = np.asarray(queries, dtype=np.float32)
        labels, distances = self.search_ids(q, top_k)
        out_k, out_s = [], []
        with self.index_lock:
            for r in range(len(q)):
                ks, ss = [], []
                for h, d in zip(labels[r].tolist(), distances[r].tolist()):
                    key = self.store.key_of(h) if h >= 0 and h not in self.deleted_ids else ""
                    if key:
                        ks.append(key)
                        ss.append(float(d))
                out_k.append(ks)
                out_s.append(ss)
        return out_k, out_s

    def get(self, key: str) -> Response:                                             # :411-428
        with self.index_lock:
            hid = self.store.id_of(key)
            if hid == -1:
                return Response(success=False, message=f"key={key}不存在")
            if hid in self.deleted_ids:
                return Response(success=False, message=f"key={key}已被删除")
            return Response(success=True, vector_data=VectorData(key=key, vector=self.store.vector(hid).tolist(),
                                                                 metadata=self.store.metadata(hid)))

    def close(self) -> None:                                                         # _on_exit, :61-72
        with self.index_lock:
            self.save_checkpoint()
            self.store.close()
            if hasattr(self.hnsw_index, "close"):
                self.hnsw_index.close()
