"""hnswlib.Index-shaped facade over the B200 exact-kNN shard.

`VectorNodeHandler` (reference src/datanode/handler.py) talks to its index only through
``hnswlib.Index`` methods; this class offers the same methods with the same argument meaning
and error behaviour (RuntimeError), so the handler switches by changing one constructor line
(handler.py:46).  Search is exact (every live row is scored), so ``set_ef`` /
``ef_construction`` / ``M`` are accepted and ignored.
"""
from __future__ import annotations

import ctypes as C
import threading
from typing import Optional, Sequence, Tuple

import numpy as np

from . import _ffi


def _as_f32_2d(x, dim: int, what: str) -> np.ndarray:
    a = np.ascontiguousarray(np.asarray(x, dtype=np.float32))
    if a.ndim == 1:
        a = a.reshape(1, -1)
    if a.ndim != 2 or a.shape[1] != dim:
        # hnswlib: "Wrong dimensionality of the vectors"
        raise RuntimeError(f"Wrong dimensionality of the vectors: {what} has shape {a.shape}, index dim is {dim}")
    return a


class Index:
    """Drop-in for ``hnswlib.Index(space, dim)``.

    Extra keyword ``store_dtype`` ('f32' | 'f16') selects the HBM storage type (the reference
    stores fp32 only); ``device`` the CUDA ordinal."""

    def __init__(self, space: str = "l2", dim: int = 512, *, store_dtype: str = "f32", device: int = 0):
        if space not in _ffi.METRIC_CODES:
            raise RuntimeError("Space name must be one of l2, ip, or cosine.")   # hnswlib's wording
        if store_dtype not in _ffi.DTYPE_CODES:
            raise RuntimeError("store_dtype must be 'f32' or 'f16'")
        self.space = space
        self.dim = int(dim)
        self.store_dtype = store_dtype
        self.device = int(device)
        self._h: Optional[int] = None
        self._ef = 10
        self._auto_label = 0
        self._lock = threading.Lock()

    # ---- lifecycle -----------------------------------------------------------------------
    def init_index(self, max_elements: int, ef_construction: int = 200, M: int = 16, random_seed: int = 100,
                   allow_replace_deleted: bool = False) -> None:
        """hnswlib.Index.init_index (handler.py:86,111).  Graph parameters are ignored."""
        if self._h is not None:
            raise RuntimeError("The index is already initiated.")
        h = C.c_void_p()
        _ffi.check(_ffi.lib().vdb_create(self.dim, _ffi.METRIC_CODES[self.space], _ffi.DTYPE_CODES[self.store_dtype],
                                         int(max_elements), self.device, C.byref(h)), "init_index")
        self._h = h.value

    def _handle(self) -> int:
        if self._h is None:
            raise RuntimeError("Search index has not been initialized, call init_index first.")
        return self._h

    def close(self) -> None:
        if self._h is not None:
            _ffi.lib().vdb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- hnswlib surface -----------------------------------------------------------------
    def set_ef(self, ef: int) -> None:
        """handler.py:88,113,361 -- no-op: the search is exact."""
        self._ef = int(ef)

    def set_num_threads(self, n: int) -> None:
        pass

    def get_current_count(self) -> int:
        """handler.py:82,196,237,350 -- appended rows, tombstoned ones included."""
        return int(_ffi.lib().vdb_count(self._handle()))

    def get_max_elements(self) -> int:
        """handler.py:238."""
        return int(_ffi.lib().vdb_capacity(self._handle()))

    def get_live_count(self) -> int:
        return int(_ffi.lib().vdb_live_count(self._handle()))

    def add_items(self, data, ids=None, num_threads: int = -1, replace_deleted: bool = False) -> None:
        """handler.py:112,268-271.  RuntimeError when the index is full, as hnswlib."""
        h = self._handle()
        rows = _as_f32_2d(data, self.dim, "data")
        n = rows.shape[0]
        if ids is None:
            with self._lock:
                ids_a = np.arange(self._auto_label, self._auto_label + n, dtype=np.int64)
        else:
            ids_a = np.ascontiguousarray(np.asarray(ids, dtype=np.int64).reshape(-1))
            if ids_a.shape[0] != n:
                raise RuntimeError("wrong dimensionality of the labels")
        _ffi.check(_ffi.lib().vdb_add(h, rows.ctypes.data_as(_ffi._f32p), ids_a.ctypes.data_as(_ffi._i64p), n),
                   "add_items")
        with self._lock:
            if n:
                self._auto_label = max(self._auto_label, int(ids_a.max()) + 1)

    def knn_query(self, data, k: int = 1, num_threads: int = -1, filter=None) -> Tuple[np.ndarray, np.ndarray]:
        """handler.py:364.  Returns (labels uint64 [nq,k], distances float32 [nq,k]) ascending.
        Like hnswlib, raises RuntimeError when fewer than k live elements exist."""
        if filter is not None:
            raise RuntimeError("filter callbacks are not supported by the exact GPU index")
        labels, dist, counts = self.knn_query_padded(data, k)
        if labels.shape[0] and int(counts.min()) < k:
            raise RuntimeError("Cannot return the results in a contiguous 2D array. Probably ef or M is too small")
        return labels.astype(np.uint64), dist

    def knn_query_padded(self, data, k: int, out=None) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """Exact search that does not fail on short results: labels int64 (-1 padded), distances
        (+inf padded), counts int32 [nq].  `out` = (labels, dist, counts) arrays to fill (e.g. page-locked
        ones from `pinned_empty`, which the library reads and writes by DMA without a staging copy)."""
        h = self._handle()
        q = _as_f32_2d(data, self.dim, "query")
        nq = q.shape[0]
        k = int(k)
        if out is not None:
            labels, dist, counts = out
            if (labels.shape != (nq, k) or dist.shape != (nq, k) or counts.shape != (nq,) or labels.dtype != np.int64
                    or dist.dtype != np.float32 or counts.dtype != np.int32
                    or not (labels.flags.c_contiguous and dist.flags.c_contiguous and counts.flags.c_contiguous)):
                raise RuntimeError("out must be C-contiguous (int64 [nq,k], float32 [nq,k], int32 [nq])")
        else:
            labels = np.empty((nq, k), dtype=np.int64)
            dist = np.empty((nq, k), dtype=np.float32)
            counts = np.empty((nq,), dtype=np.int32)
        _ffi.check(_ffi.lib().vdb_search(h, q.ctypes.data_as(_ffi._f32p), nq, k, labels.ctypes.data_as(_ffi._i64p),
                                         dist.ctypes.data_as(_ffi._f32p), counts.ctypes.data_as(_ffi._i32p)),
                   "knn_query")
        return labels, dist, counts

    def submit_query(self, data, k: int, out=None):
        """First half of `knn_query_padded`: enqueue upload + search + download and return a ticket for
        `collect_query`.  A caller that keeps two tickets open overlaps the upload of one batch with the search
        of the other (vdb_search_submit / vdb_search_collect).  `data` and `out` must stay untouched until
        the ticket has been collected."""
        h = self._handle()
        q = _as_f32_2d(data, self.dim, "query")
        nq = q.shape[0]
        k = int(k)
        if out is not None:
            labels, dist, counts = out
            if (labels.shape != (nq, k) or dist.shape != (nq, k) or counts.shape != (nq,) or labels.dtype != np.int64
                    or dist.dtype != np.float32 or counts.dtype != np.int32
                    or not (labels.flags.c_contiguous and dist.flags.c_contiguous and counts.flags.c_contiguous)):
                raise RuntimeError("out must be C-contiguous (int64 [nq,k], float32 [nq,k], int32 [nq])")
        else:
            labels = np.empty((nq, k), dtype=np.int64)
            dist = np.empty((nq, k), dtype=np.float32)
            counts = np.empty((nq,), dtype=np.int32)
        t = _ffi.C.c_void_p()
        _ffi.check(_ffi.lib().vdb_search_submit(h, q.ctypes.data_as(_ffi._f32p), nq, k, labels.ctypes.data_as(_ffi._i64p),
                                                dist.ctypes.data_as(_ffi._f32p), counts.ctypes.data_as(_ffi._i32p),
                                                _ffi.C.byref(t)), "submit_query")
        return (t, q, labels, dist, counts)          # keeps the buffers alive

    def collect_query(self, ticket) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        t, _q, labels, dist, counts = ticket
        _ffi.check(_ffi.lib().vdb_search_collect(t), "collect_query")
        return labels, dist, counts

    def mark_deleted(self, label) -> None:
        """hnswlib.Index.mark_deleted; replaces the handler's python-side filter (handler.py:378)."""
        ids = np.ascontiguousarray(np.asarray(label, dtype=np.int64).reshape(-1))
        _ffi.check(_ffi.lib().vdb_mark_deleted(self._handle(), ids.ctypes.data_as(_ffi._i64p), ids.shape[0]),
                   "mark_deleted")

    def unmark_deleted(self, label) -> None:
        ids = np.ascontiguousarray(np.asarray(label, dtype=np.int64).reshape(-1))
        _ffi.check(_ffi.lib().vdb_unmark_deleted(self._handle(), ids.ctypes.data_as(_ffi._i64p), ids.shape[0]),
                   "unmark_deleted")

    def get_items(self, ids, return_type: str = "numpy"):
        ids_a = np.ascontiguousarray(np.asarray(ids, dtype=np.int64).reshape(-1))
        out = np.empty((ids_a.shape[0], self.dim), dtype=np.float32)
        _ffi.check(_ffi.lib().vdb_get_rows(self._handle(), ids_a.ctypes.data_as(_ffi._i64p), ids_a.shape[0],
                                           out.ctypes.data_as(_ffi._f32p)), "get_items")
        return out if return_type == "numpy" else out.tolist()

    def resize_index(self, new_size: int) -> None:
        _ffi.check(_ffi.lib().vdb_resize(self._handle(), int(new_size)), "resize_index")

    def save_index(self, path: str) -> None:
        """handler.py:65,115,164,302."""
        _ffi.check(_ffi.lib().vdb_save(self._handle(), str(path).encode()), "save_index")

    def load_index(self, path: str, max_elements: int = 0, allow_replace_deleted: bool = False) -> None:
        """handler.py:80,195."""
        h = C.c_void_p()
        _ffi.check(_ffi.lib().vdb_load(str(path).encode(), int(max_elements), self.device, C.byref(h)), "load_index")
        got_dim = int(_ffi.lib().vdb_dim(h.value))
        if got_dim != self.dim:
            _ffi.lib().vdb_destroy(h.value)
            raise RuntimeError(f"snapshot dim {got_dim} does not match index dim {self.dim}")
        self.close()
        self._h = h.value
        self._auto_label = self.get_current_count()

    def save_image(self, image_dir: str, meta_path: str) -> None:
        """Incremental form of save_index for frequent checkpoints: the shard's append-only image under `image_dir`
        grows by the rows added since the last call; `meta_path` gets the small header + tombstone bitmap of this
        moment (what a checkpoint stores as its index.bin)."""
        _ffi.check(_ffi.lib().vdb_save_image(self._handle(), str(image_dir).encode(), str(meta_path).encode()), "save_image")

    def load_image(self, image_dir: str, meta_path: str, max_elements: int = 0) -> None:
        h = C.c_void_p()
        _ffi.check(_ffi.lib().vdb_load_image(str(image_dir).encode(), str(meta_path).encode(), int(max_elements), self.device,
                                             C.byref(h)), "load_image")
        got_dim = int(_ffi.lib().vdb_dim(h.value))
        if got_dim != self.dim:
            _ffi.lib().vdb_destroy(h.value)
            raise RuntimeError(f"image dim {got_dim} does not match index dim {self.dim}")
        self.close()
        self._h = h.value
        self._auto_label = self.get_current_count()

    @staticmethod
    def is_image_meta(path: str) -> bool:
        try:
            with open(path, "rb") as f:
                return f.read(8) == b"VDBIMG2\0"
        except OSError:
            return False

    # ---- extensions used by benches / the sharded path ----------------------------------
    def add_synthetic(self, seed: int, row_start: int, n: int, label_start: Optional[int] = None) -> None:
        """Append rows of the synthetic unit-norm set generated on the device (bench utility)."""
        ls = row_start if label_start is None else label_start
        _ffi.check(_ffi.lib().vdb_add_synthetic(self._handle(), seed, row_start, n, ls), "add_synthetic")
        with self._lock:
            self._auto_label = max(self._auto_label, ls + n)

    def search_device(self, d_queries_ptr: int, nq: int, k: int, d_labels_ptr: int, d_dist_ptr: int,
                      d_counts_ptr: int = 0, stream: int = 0) -> None:
        """Enqueue a search whose inputs/outputs are device pointers (torch ``.data_ptr()``)."""
        _ffi.check(_ffi.lib().vdb_search_dev(self._handle(), d_queries_ptr, nq, int(k), d_labels_ptr, d_dist_ptr,
                                             d_counts_ptr or None, stream or None), "search_device")

    def set_option(self, name: str, value: int) -> None:
        _ffi.check(_ffi.lib().vdb_set_option(self._handle(), name.encode(), int(value)), "set_option")

    def get_stat(self, name: str) -> int:
        return int(_ffi.lib().vdb_get_stat(self._handle(), name.encode()))


class _PinnedBlock:
    def __init__(self, nbytes: int):
        self.ptr = _ffi.lib().vdb_host_alloc(nbytes)
        if not self.ptr:
            raise RuntimeError(_ffi.last_error())

    def __del__(self):
        try:
            _ffi.lib().vdb_host_free(self.ptr)
        except Exception:
            pass


def pinned_empty(shape, dtype=np.float32) -> np.ndarray:
    """numpy array over page-locked host memory (vdb_host_alloc): query / result buffers that
    `Index.knn_query_padded` moves by DMA with no staging copy.  Freed with the array."""
    dt = np.dtype(dtype)
    n = int(np.prod(shape)) if np.ndim(shape) else int(shape)
    blk = _PinnedBlock(max(n * dt.itemsize, 1))
    buf = (C.c_ubyte * max(n * dt.itemsize, 1)).from_address(blk.ptr)
    arr = np.frombuffer(buf, dtype=dt, count=n).reshape(shape)
    _keep_alive[id(buf)] = blk
    import weakref
    weakref.finalize(buf, _keep_alive.pop, id(buf), None)
    return arr


_keep_alive = {}


def merge_topk(dist: np.ndarray, ids: np.ndarray, k: int, device: int = 0) -> Tuple[np.ndarray, np.ndarray]:
    """Host-buffer form of the cross-shard merge (coordinator/handler.py:212-216) run on the GPU:
    dist/ids [G, nq, k_in] -> (dist [nq,k], ids [nq,k]) ascending (distance, id)."""
    dist = np.ascontiguousarray(dist, dtype=np.float32)
    ids = np.ascontiguousarray(ids, dtype=np.int64)
    g, nq, k_in = dist.shape
    o_d = np.empty((nq, k), dtype=np.float32)
    o_i = np.empty((nq, k), dtype=np.int64)
    _ffi.check(_ffi.lib().vdb_merge_topk(dist.ctypes.data, ids.ctypes.data, g, nq, k_in, k, o_d.ctypes.data,
                                         o_i.ctypes.data, 0, device, None), "merge_topk")
    return o_d, o_i


def launch_count() -> int:
    return int(_ffi.lib().vdb_launch_count())
