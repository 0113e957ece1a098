"""Coordinator side of the search path: scatter to every shard, gather, merge
(reference src/coordinator/handler.py:117-228).  RPC pooling and ZooKeeper are out of scope; the
"nodes" here are handler objects living in this process (one per GPU), or ranks of a
torch.distributed job (`ShardedSearcher`)."""
from __future__ import annotations

import os

from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

from .sharding import assign_shards_to_nodes, get_shard_id
from .ttypes import Response, SearchRequest, SearchResult, VectorData


def merge_search_results(per_node: Sequence[Optional[SearchResult]], top_k: int) -> SearchResult:
    """The merge of `CoordinatorHandler.search` (coordinator/handler.py:200-225): concatenate in node order,
    the first occurrence of a key wins, stable ascending sort by score, slice top_k."""
    all_keys, all_scores, all_vectors, seen = [], [], [], set()
    for res in per_node:
        if res is None:
            continue
        vecs = res.vectors if res.vectors is not None else [None] * len(res.keys or [])
        for k, s, v in zip(res.keys or [], res.scores or [], vecs):
            if k in seen:
                continue
            seen.add(k)
            all_keys.append(k)
            all_scores.append(s)
            all_vectors.append(v)
    if not all_scores:
        return SearchResult([], [], [])
    order = sorted(range(len(all_scores)), key=lambda i: all_scores[i])[:top_k]
    return SearchResult(keys=[all_keys[i] for i in order], scores=[all_scores[i] for i in order],
                        vectors=[all_vectors[i] for i in order])


class LocalCoordinator:
    """`CoordinatorHandler` over handlers in this process (one per GPU): md5 routing of put/delete/get
    to the shard master (coordinator/handler.py:117-170), broadcast search + merge (:173-228)."""

    def __init__(self, nodes: Dict[str, object], shard_count: Optional[int] = None, parallel: bool = True):
        self.nodes = dict(nodes)
        self._pool = None
        if parallel and len(nodes) > 1:
            from concurrent.futures import ThreadPoolExecutor
            self._pool = ThreadPoolExecutor(max_workers=len(nodes), thread_name_prefix="coord-fanout")
        self.node_ids = list(self.nodes)
        # with SHARD_COUNT=4 and 8 nodes only 4 nodes would receive data: default to one shard per node
        self.shard_count = shard_count or len(self.node_ids)
        self.mapping = assign_shards_to_nodes(self.node_ids, self.shard_count)

    def _master(self, key: str):
        return self.nodes[self.mapping[get_shard_id(key, self.shard_count)]["master"]]

    def put(self, data: VectorData) -> Response:
        return self._master(data.key).put(data)

    def delete(self, key: str) -> Response:
        return self._master(key).delete(key)

    def get(self, key: str) -> Response:
        return self._master(key).get(key)

    def search(self, req: SearchRequest) -> Response:
        if not self.nodes:
            return Response(success=False, message="无在线数据节点")               # :177-178
        sub = SearchRequest(query_vector=req.query_vector, top_k=req.top_k)       # :186-189 (filter/threshold dropped)

        def one(node_id):
            try:
                resp = self.nodes[node_id].search(sub)
            except Exception:                                                      # :198-199 a failed node is skipped
                return None
            return resp.search_result if resp.success and resp.search_result else None

        if self._pool is not None:
            # the reference asks the nodes one after the other (:191-197); here all at once -- the GPU calls drop the
            # GIL -- and the answers are merged in NODE order, so ties resolve exactly as in the serial loop
            results = list(self._pool.map(one, self.node_ids))
        else:
            results = [one(node_id) for node_id in self.node_ids]                  # :191
        merged = merge_search_results(results, req.top_k)
        return Response(success=True, search_result=merged)

    NODE_SHIFT = 28          # merged labels carry (node index << 28 | node-local id): 16 nodes x 268M rows

    def search_batch(self, queries, top_k: int, gpu_merge: Optional[bool] = None) -> Tuple[List[List[str]], List[List[float]]]:
        """Additive: many queries per call.  Every node answers with ARRAYS (`search_ids`: the tensor-core path,
        tombstones masked on the GPU), the G lists of each query are merged on the GPU (`vdb_merge_topk`, kernel K5:
        ascending (score, node, id) -- the order of the reference's stable sort over the node-ordered concatenation,
        coordinator/handler.py:200-216) and only the k winners of each query are mapped to keys.  Nodes without
        `search_ids` (remote / plain handlers), or `gpu_merge=False`, take the per-query Python merge.
        Returns (keys[nq][<=k], scores[nq][<=k])."""
        nq = len(queries)
        if not self.nodes:
            return [[] for _ in range(nq)], [[] for _ in range(nq)]
        k = top_k if top_k and top_k > 0 else 5
        array_nodes = all(hasattr(n, "search_ids") and hasattr(n, "keys_of") for n in self.nodes.values())
        if gpu_merge is None:
            gpu_merge = array_nodes
        if gpu_merge and array_nodes and len(self.node_ids) <= (1 << (32 - self.NODE_SHIFT)):
            return self._search_batch_gpu(np.ascontiguousarray(np.asarray(queries, dtype=np.float32)), k)

        def one(node_id):
            try:
                return self.nodes[node_id].search_batch(queries, k)
            except Exception:
                return None

        per_node = list(self._pool.map(one, self.node_ids)) if self._pool is not None else [one(n) for n in self.node_ids]
        out_k, out_s = [], []
        for r in range(nq):
            lists = [SearchResult(keys=pn[0][r], scores=pn[1][r], vectors=None) for pn in per_node if pn is not None]
            m = merge_search_results(lists, k)
            out_k.append(m.keys)
            out_s.append(m.scores)
        return out_k, out_s

    def _search_batch_gpu(self, q: np.ndarray, k: int):
        from .index import merge_topk
        nq = len(q)

        def one(node_id):
            try:
                return self.nodes[node_id].search_ids(q, k)
            except Exception:                                                      # :198-199 a failed node is skipped
                return None

        per_node = list(self._pool.map(one, self.node_ids)) if self._pool is not None else [one(n) for n in self.node_ids]
        live = [(g, pn) for g, pn in enumerate(per_node) if pn is not None]
        if not live:
            return [[] for _ in range(nq)], [[] for _ in range(nq)]
        mask = (1 << self.NODE_SHIFT) - 1
        ids = np.stack([np.where(pn[0] >= 0, pn[0] + (g << self.NODE_SHIFT), -1) for g, pn in live]).astype(np.int64)
        if any(int(pn[0].max(initial=-1)) > mask for _, pn in live):
            raise RuntimeError("node-local ids beyond 2^28: merge on the host instead (gpu_merge=False)")
        dist = np.stack([pn[1] for _, pn in live]).astype(np.float32)
        device = getattr(self.nodes[self.node_ids[live[0][0]]], "device", 0)
        m_d, m_i = merge_topk(dist, ids, k, device=device)                        # [nq, k] ascending (score, node, id)
        out_k, out_s = [], []
        # ids -> keys once per node for all winners
        node_of = np.where(m_i >= 0, m_i >> self.NODE_SHIFT, -1)
        local = np.where(m_i >= 0, m_i & mask, -1)
        key_grid = np.full(m_i.shape, "", dtype=object)
        for g, _ in live:
            sel = node_of == g
            if sel.any():
                key_grid[sel] = self.nodes[self.node_ids[g]].keys_of(local[sel].tolist())
        for r in range(nq):
            ks, ss, seen = [], [], set()
            for key, d in zip(key_grid[r].tolist(), m_d[r].tolist()):
                if key and key not in seen:                                        # first seen wins (:200-206)
                    seen.add(key)
                    ks.append(key)
                    ss.append(float(d))
            out_k.append(ks)
            out_s.append(ss)
        return out_k, out_s

    def close(self) -> None:
        if self._pool is not None:
            self._pool.shutdown(wait=False)
            self._pool = None


class ShardedSearcher:
    """One rank per GPU (torch.distributed): every rank searches its shard for the whole batch; the per-rank
    top-k lists are exchanged BY QUERY SLICE (all-to-all: rank r receives every rank's lists for queries
    [r*nq/G, (r+1)*nq/G)) and merged by (distance, id) on the owner of the slice -- the GPU form of
    coordinator/handler.py:191-216 with the merge work and the exchanged bytes divided by G.  NCCL over NVLink
    on GPUs; gloo in the CPU tests.  Batches that do not divide by G (a single query) are all-gathered and
    merged on every rank.

    `local_search(queries[nq, dim], k) -> (ids int64 [nq,k] global ids, -1 padded; dist float32 [nq,k])`
    `merge(dist [G,n,k], ids [G,n,k], k) -> (dist [n,k], ids [n,k])`   (the CUDA merge kernel in production)
    Tensors are torch tensors on the device of the process group's backend."""

    def __init__(self, local_search: Callable, merge: Callable, group=None):
        import torch.distributed as dist
        self._dist = dist
        self.local_search, self.merge, self.group = local_search, merge, group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)

    def search_slice(self, queries, k: int):
        """Results of this rank's query slice only: (dist [nq/G,k], ids [nq/G,k]); nq must divide by G."""
        import torch
        ids, dd = self.local_search(queries, k)
        nq = ids.shape[0]
        if nq % self.world:
            raise ValueError("search_slice needs a batch that divides by the world size")
        g_ids = torch.empty_like(ids)
        g_dd = torch.empty_like(dd)
        self._dist.all_to_all_single(g_ids, ids.contiguous(), group=self.group)
        self._dist.all_to_all_single(g_dd, dd.contiguous(), group=self.group)
        sl = nq // self.world
        return self.merge(g_dd.view(self.world, sl, -1), g_ids.view(self.world, sl, -1), k)

    def search(self, queries, k: int):
        """Full result on every rank: (dist [nq,k], ids [nq,k])."""
        import torch
        nq = queries.shape[0]
        if nq % self.world == 0 and nq >= self.world:
            dd, ids = self.search_slice(queries, k)
            o_dd = torch.empty((nq,) + tuple(dd.shape[1:]), dtype=dd.dtype, device=dd.device)
            o_ids = torch.empty((nq,) + tuple(ids.shape[1:]), dtype=ids.dtype, device=ids.device)
            self._dist.all_gather_into_tensor(o_dd, dd.contiguous(), group=self.group)
            self._dist.all_gather_into_tensor(o_ids, ids.contiguous(), group=self.group)
            return o_dd, o_ids
        ids, dd = self.local_search(queries, k)
        # concatenated along dim 0 (the layout both NCCL and gloo accept), viewed as [G, nq, k]
        g_ids = torch.empty((self.world * nq,) + tuple(ids.shape[1:]), dtype=ids.dtype, device=ids.device)
        g_dd = torch.empty((self.world * nq,) + tuple(dd.shape[1:]), dtype=dd.dtype, device=dd.device)
        self._dist.all_gather_into_tensor(g_ids, ids.contiguous(), group=self.group)
        self._dist.all_gather_into_tensor(g_dd, dd.contiguous(), group=self.group)
        return self.merge(g_dd.view(self.world, nq, -1), g_ids.view(self.world, nq, -1), k)


class PeerExchange:
    """Exchange + merge of the per-rank top-k lists in ONE CUDA kernel over NVLink peer memory (C ABI
    vdb_xchg_*, kernel K5x in csrc/merge_topk.cu): the NCCL-free form of `ShardedSearcher.search_slice`'s
    all-to-all + merge for the ranks of one box.

    `exchange_handles(my_handle: bytes) -> list[bytes]` must return every rank's 64-byte handle in rank order
    (e.g. an all-gather over the job's process group)."""

    def __init__(self, device: int, rank: int, world: int, max_slice: int, max_k: int, exchange_handles: Callable,
                 query_slot_bytes: int = 0):
        import ctypes as C
        from . import _ffi
        self._ffi, self._C = _ffi, C
        self.rank, self.world = rank, world
        self.query_slot_bytes = int(query_slot_bytes)
        h = C.c_void_p()
        buf = (C.c_ubyte * 64)()
        _ffi.check(_ffi.lib().vdb_xchg_create_q(device, rank, world, max_slice, max_k, self.query_slot_bytes, C.byref(h), buf),
                   "xchg_create")
        self._h = h.value
        handles = exchange_handles(bytes(buf))
        if len(handles) != world or any(len(x) != 64 for x in handles):
            raise RuntimeError("exchange_handles must return one 64-byte handle per rank")
        blob = (C.c_ubyte * (64 * world)).from_buffer_copy(b"".join(handles))
        _ffi.check(_ffi.lib().vdb_xchg_connect(self._h, blob), "xchg_connect")

    def merge(self, d_dist_ptr: int, d_ids_ptr: int, nq: int, k: int, o_dist_ptr: int, o_ids_ptr: int, stream: int = 0):
        """Enqueue on `stream`: this rank's lists [nq,k] (device pointers) -> its slice's results [nq/world,k]."""
        self._ffi.check(self._ffi.lib().vdb_xchg_merge_dev(self._h, d_dist_ptr, d_ids_ptr, nq, int(k), o_dist_ptr,
                                                           o_ids_ptr, stream or None), "xchg_merge")

    # ---- query all-gather over the copy engines (needs query_slot_bytes > 0) -----------------------------------
    def query_slot(self, slot: int) -> int:
        """device pointer of local query slot 0 / 1: the whole batch, [world][slice] rows in rank order"""
        return int(self._ffi.lib().vdb_xchg_query_slot(self._h, int(slot)) or 0)

    def gather_queries(self, slice_ptr: int, slice_bytes: int, slot: int, batch_no: int, stream: int) -> None:
        """Enqueue on `stream` (a copy stream): this rank's slice (host or device pointer) -> its place in slot `slot`
        here and, by DMA over NVLink, on every peer; then the arrival word `batch_no` in every rank's array."""
        self._ffi.check(self._ffi.lib().vdb_xchg_gather_queries(self._h, slice_ptr, slice_bytes, int(slot), int(batch_no),
                                                                stream or None), "xchg_gather_queries")

    def wait_queries(self, slot: int, batch_no: int, stream: int) -> None:
        """Enqueue on `stream` (the search stream): wait until every rank's slice of `batch_no` is in the local slot."""
        self._ffi.check(self._ffi.lib().vdb_xchg_wait_queries(self._h, int(slot), int(batch_no), stream or None),
                        "xchg_wait_queries")

    def status(self) -> None:
        """Raises RuntimeError if a step failed (a peer never arrived, or arrived with another batch size / k).
        Call after synchronising the stream the step was enqueued on."""
        self._ffi.check(self._ffi.lib().vdb_xchg_status(self._h), "xchg_status")

    def close(self):
        if self._h is not None:
            self._ffi.lib().vdb_xchg_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ShardedIndex:
    """One rank of a database sharded by rows over the GPUs of one box (one process per GPU in a torch.distributed
    job; GPU g holds what datanode g would): the call a client of rank r makes -- ITS SLICE of the query batch in,
    the merged results of that slice out -- i.e. coordinator/handler.py:186-216 with every step on the GPUs:

      1. each rank uploads its slice of the queries; the slices are all-gathered over NVLink (every shard must see
         the whole batch -- the broadcast of :191-197)
      2. each rank searches its shard for the whole batch (`Index.search_device`: K1 / K2)
      3. the per-rank lists are exchanged by query slice and merged on the slice's owner (:200-216): the fused
         NVLink exchange + merge kernel (`exchange="p2p"`, K5x) or NCCL all-to-all + the merge kernel ("nccl")

    A batch that does not divide by the world size (a single query) is replicated: every rank passes the whole
    batch and gets the whole result.  Buffers are allocated once for `max_batch` x `max_k`."""

    def __init__(self, index, *, max_batch: int, max_k: int, exchange: str = "p2p", group=None):
        import torch
        import torch.distributed as dist
        from . import _ffi
        self._torch, self._dist, self._ffi = torch, dist, _ffi
        self.ix, self.group = index, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.dev = torch.device("cuda", index.device)
        self.max_batch, self.max_k = int(max_batch), int(max_k)
        sl = (self.max_batch + self.world - 1) // self.world
        self.px = None
        if exchange == "p2p" and self.world > 1:
            self.px = PeerExchange(index.device, self.rank, self.world, max_slice=sl, max_k=self.max_k,
                                   exchange_handles=self._gather_handles,
                                   query_slot_bytes=sl * self.world * index.dim * 4)
        elif exchange not in ("p2p", "nccl"):
            raise ValueError("exchange must be 'p2p' or 'nccl'")
        e = lambda *shape, dtype: torch.empty(shape, dtype=dtype, device=self.dev)          # noqa: E731
        self._q = e(self.max_batch, index.dim, dtype=torch.float32)
        self._ids = e(self.max_batch * self.max_k, dtype=torch.int64)
        self._dd = e(self.max_batch * self.max_k, dtype=torch.float32)
        self._g_ids = e(self.world * self.max_batch * self.max_k, dtype=torch.int64)         # nccl path / replicated batch
        self._g_dd = e(self.world * self.max_batch * self.max_k, dtype=torch.float32)
        self._o_ids = e(self.max_batch * self.max_k, dtype=torch.int64)
        self._o_dd = e(self.max_batch * self.max_k, dtype=torch.float32)
        self._views = {}
        self._ce_gather = os.environ.get("VDB_CE_GATHER", "1") != "0"      # A/B switch of the copy-engine query gather

    def close(self) -> None:
        """Releases the peer-memory exchange (collective: every rank closes before the process group goes away)."""
        if self.px is not None:
            self.px.close()
            self.px = None

    def _gather_handles(self, mine: bytes):
        torch, dist = self._torch, self._dist
        t = torch.frombuffer(bytearray(mine), dtype=torch.uint8).to(self.dev)
        o = torch.empty((self.world, 64), dtype=torch.uint8, device=self.dev)
        dist.all_gather_into_tensor(o, t, group=self.group)
        return [o[r].cpu().numpy().tobytes() for r in range(self.world)]

    def slice_of(self, nq: int) -> Tuple[int, int]:
        """rows [lo, hi) of a batch of nq queries whose results this rank gets back: slices of ceil(nq / world)
        (a single query belongs to rank 0; ranks past the end own nothing)"""
        s = (nq + self.world - 1) // self.world
        return min(nq, self.rank * s), min(nq, (self.rank + 1) * s)

    def search_device(self, d_queries, k: int):
        """`d_queries` [nq, dim] fp32 on this rank's GPU, the WHOLE batch -> (dist [n, k], ids [n, k]) device tensors
        of this rank's slice (`slice_of(nq)`; views of internal buffers, valid until the next call; enqueued on the
        current stream)."""
        return self._search_ptr(d_queries.data_ptr(), int(d_queries.shape[0]), k)

    def _search_ptr(self, q_ptr: int, nq: int, k: int):
        torch, dist = self._torch, self._dist
        if nq > self.max_batch or k > self.max_k:
            raise ValueError("batch or k beyond what this ShardedIndex was sized for")
        stream = torch.cuda.current_stream().cuda_stream
        v = self._views.get((nq, k))
        if v is None:                 # tensor views per batch shape, made once (they cost more than the launches)
            lo, hi = self.slice_of(nq)
            s = (nq + self.world - 1) // self.world
            v = self._views[(nq, k)] = (self._ids[:nq * k].view(nq, k), self._dd[:nq * k].view(nq, k), lo, hi, s,
                                        self._o_ids[:s * k].view(s, k), self._o_dd[:s * k].view(s, k),
                                        self._o_ids[:s * k].view(s, k)[:hi - lo], self._o_dd[:s * k].view(s, k)[:hi - lo])
        ids, dd, lo, hi, s, o_ids, o_dd, r_ids, r_dd = v
        self.ix.search_device(q_ptr, nq, k, ids.data_ptr(), dd.data_ptr(), 0, stream)
        if self.world == 1:
            return dd, ids
        if self.px is not None:       # one kernel: peer stores over NVLink, flags, merge of the owned slice (ragged too)
            self.px.merge(dd.data_ptr(), ids.data_ptr(), nq, k, o_dd.data_ptr(), o_ids.data_ptr(), stream)
            return r_dd, r_ids
        lib = self._ffi.lib()
        if nq % self.world == 0:      # rank r receives every rank's lists for slice r
            g_ids, g_dd = self._g_ids[:nq * k].view(nq, k), self._g_dd[:nq * k].view(nq, k)
            dist.all_to_all_single(g_ids, ids, group=self.group)
            dist.all_to_all_single(g_dd, dd, group=self.group)
            self._ffi.check(lib.vdb_merge_topk(g_dd.data_ptr(), g_ids.data_ptr(), self.world, s, k, k, o_dd.data_ptr(),
                                               o_ids.data_ptr(), 1, self.ix.device, stream), "merge")
            return o_dd, o_ids
        # ragged batch: all-gather the lists, every rank merges, keeps its slice
        g_ids = self._g_ids[:self.world * nq * k].view(self.world * nq, k)
        g_dd = self._g_dd[:self.world * nq * k].view(self.world * nq, k)
        dist.all_gather_into_tensor(g_ids, ids, group=self.group)
        dist.all_gather_into_tensor(g_dd, dd, group=self.group)
        a_ids, a_dd = self._ids[:nq * k].view(nq, k), self._dd[:nq * k].view(nq, k)      # the local lists are spent
        self._ffi.check(lib.vdb_merge_topk(g_dd.data_ptr(), g_ids.data_ptr(), self.world, nq, k, k, a_dd.data_ptr(),
                                           a_ids.data_ptr(), 1, self.ix.device, stream), "merge")
        return a_dd[lo:hi], a_ids[lo:hi]

    # ---- pipelined client calls: upload of batch i+1 overlaps the search of batch i ---------------------------
    def submit_host(self, q, k: int):
        """Asynchronous `search_host` for even batches (every rank passes ITS rows of the batch, page-locked): the
        upload of the queries runs on a copy stream (DMA) into one of two query buffers; the NVLink all-gather of the
        slices, the search, the exchange + merge and the download of the results are enqueued on the current stream
        behind it.  Returns a ticket for `collect`.  With two batches in flight the host<->device copies of one batch
        (and the host's launch work) hide behind the search of the other; at most two tickets may be outstanding.

        With the peer-memory exchange (`exchange="p2p"`) the all-gather of the slices runs on the COPY ENGINES too
        (`PeerExchange.gather_queries`: the upload, then one DMA per peer over NVLink into the peers' query slots, then a
        one-warp kernel that publishes the batch number; the search stream only waits for the arrival words), so it is
        hidden behind the previous batch as well.  With `exchange="nccl"` the all-gather is an NCCL kernel and stays on
        the SEARCH stream on purpose: it is a kernel that waits for its peers, and so is the fused exchange + merge.
        Two such kernels in flight at once on one GPU can deadlock a box (rank A: the all-gather is resident and keeps
        the cooperative exchange kernel from being placed; rank B: the exchange kernel is resident, spins for A, and
        its registers keep B's all-gather from being placed -- seen on 8 GPUs, caught by the exchange's bounded wait).
        The copy-engine form has no such kernel: the signal waits for nothing, the wait is one warp in front of the
        search on the same stream."""
        torch, dist = self._torch, self._dist
        if not torch.is_tensor(q):
            q = torch.from_numpy(q)
        n = int(q.shape[0])
        if n * self.world > self.max_batch:
            raise ValueError("batch beyond what this ShardedIndex was sized for")
        if not hasattr(self, "_pipe"):
            e = lambda *shape, dtype: torch.empty(shape, dtype=dtype, device=self.dev)          # noqa: E731
            self._pipe = {"stream": torch.cuda.Stream(device=self.dev), "slot": 0,
                          "q": [self._q, e(self.max_batch, self.ix.dim, dtype=torch.float32)],
                          "free": [torch.cuda.Event(), torch.cuda.Event()], "out": [None, None]}
        P = self._pipe
        cur = torch.cuda.current_stream()
        up = P["stream"]
        if self.px is not None and self.px.query_slot_bytes and self._ce_gather:
            if not (q.is_pinned() and q.is_contiguous() and q.dtype == torch.float32):
                raise ValueError("submit_host needs a contiguous float32 tensor in page-locked host memory")
            batch_no = P["batch_no"] = P.get("batch_no", 0) + 1
            slot = batch_no & 1
            # this rank's merge of batch_no - 2 has completed => every peer has finished searching that batch, i.e.
            # reading this slot: the DMA into the peers' slots may start
            up.wait_event(P["free"][slot])
            self.px.gather_queries(q.data_ptr(), n * self.ix.dim * 4, slot, batch_no, up.cuda_stream)
            self.px.wait_queries(slot, batch_no, cur.cuda_stream)
            dd, ids = self._search_ptr(self.px.query_slot(slot), n * self.world, k)
            return self._finish_submit(P, slot, cur, dd, ids, q)
        slot = P["slot"]
        P["slot"] ^= 1
        up.wait_event(P["free"][slot])                     # the search that last read this query buffer has finished
        d_q = P["q"][slot][:n * self.world]
        mine = d_q[self.rank * n:(self.rank + 1) * n]
        with torch.cuda.stream(up):
            mine.copy_(q, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(up)
        cur.wait_event(ready)
        if self.world > 1:
            dist.all_gather_into_tensor(d_q, mine, group=self.group)     # on the search stream: see above
        dd, ids = self.search_device(d_q, k)
        return self._finish_submit(P, slot, cur, dd, ids, None)

    def _finish_submit(self, P, slot, cur, dd, ids, keep):
        torch = self._torch
        P["free"][slot].record(cur)
        P["keep"] = (P.get("keep", (None, None))[1], keep)      # the host queries of the two batches in flight stay alive
        out = P["out"][slot]
        if out is None or tuple(out[0].shape) != tuple(ids.shape):
            out = P["out"][slot] = (torch.empty(tuple(ids.shape), dtype=torch.int64).pin_memory(),
                                    torch.empty(tuple(dd.shape), dtype=torch.float32).pin_memory())
        out[0].copy_(ids, non_blocking=True)
        out[1].copy_(dd, non_blocking=True)
        done = torch.cuda.Event()
        done.record(cur)
        return (done, out)

    def collect(self, ticket):
        """Waits for a `submit_host` ticket: (ids int64 [m, k], dist float32 [m, k]) page-locked host tensors with the
        results of the rows this rank owns; valid until the ticket after next is submitted."""
        done, out = ticket
        done.synchronize()
        return out

    def search_host(self, q, k: int, *, whole_batch: bool = False, out=None):
        """The client call.  `q`: float32 [n, dim] in page-locked host memory (torch pinned tensor; a numpy array is
        wrapped) -- this rank's rows of the batch (every rank passes the same number of rows; the batch is their
        concatenation in rank order), or with `whole_batch=True` the entire batch on every rank (batches that do
        not divide by the world size: a single query).  Returns (ids int64 [m, k], dist float32 [m, k]): page-locked
        host tensors with the results of the rows this rank owns (`slice_of`), complete when the call returns.
        `out=(ids, dist)` reuses result buffers of the right shape."""
        torch, dist = self._torch, self._dist
        if not torch.is_tensor(q):
            q = torch.from_numpy(q)
        n = int(q.shape[0])
        if whole_batch or self.world == 1:
            d_q = self._q[:n]
            d_q.copy_(q, non_blocking=True)
        else:
            d_q = self._q[:n * self.world]
            mine = d_q[self.rank * n:(self.rank + 1) * n]
            mine.copy_(q, non_blocking=True)
            dist.all_gather_into_tensor(d_q, mine, group=self.group)     # in place: slice r of d_q is rank r's upload
        dd, ids = self.search_device(d_q, k)
        if out is None or tuple(out[0].shape) != tuple(ids.shape):
            out = (torch.empty(tuple(ids.shape), dtype=torch.int64).pin_memory(),
                   torch.empty(tuple(dd.shape), dtype=torch.float32).pin_memory())
        out[0].copy_(ids, non_blocking=True)
        out[1].copy_(dd, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return out
