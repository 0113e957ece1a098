"""Key / metadata / raw-vector store of one datanode.

Replaces what the reference keeps in LevelDB (src/datanode/handler.py:288-297: one JSON value
``{hnsw_id, vector, metadata}`` per key) and its O(N) reverse lookup (``_get_key_by_hnsw_id``,
handler.py:145-153: a full iterator scan with a JSON decode per entry, per search result):

  ``<node>/leveldb_data/vectors.f32``   raw fp32 vectors, row = hnsw_id, append-only (pwrite / pread)
  ``<node>/leveldb_data/keys.log``      JSON lines, append-only:  {"i": id, "k": key, "m": metadata}  (put)
                                                                   {"d": key}                          (delete)

In memory there are two dicts (key -> id, id -> key) and the non-empty metadata dicts: ~200 bytes per vector
instead of a 512-element Python list (16 KB).  Vectors are read back from the file (page cache) when a search
result or a ``get`` needs them, exactly as given at ``put`` time -- the shard itself holds the *stored* form
(normalised for cosine, fp16-rounded for fp16 shards), which is not what the reference returns.

Both files only ever grow, so a checkpoint is a pair of numbers (``position()``); recovery truncates the log
back to it (``rollback``) and lets the WAL replay re-apply what came after.  plyvel is not installable here;
an append-only file + dict is the option SURVEY.md section 8f-2 names.
"""
from __future__ import annotations

import json
import os
from typing import Dict, Iterable, List, Optional, Sequence

import numpy as np


class KeyStore:
    VEC_FILE, LOG_FILE = "vectors.f32", "keys.log"

    def __init__(self, directory: str, dim: int, fsync: bool = True):
        self.dir, self.dim, self.fsync = directory, int(dim), fsync
        self.row_bytes = self.dim * 4
        os.makedirs(directory, exist_ok=True)
        self._fd = os.open(os.path.join(directory, self.VEC_FILE), os.O_RDWR | os.O_CREAT, 0o644)
        self._log_path = os.path.join(directory, self.LOG_FILE)
        self._id_of: Dict[str, int] = {}
        self._key_of: Dict[int, str] = {}
        self._meta: Dict[int, dict] = {}
        self._log_bytes = 0
        self._load()
        self._log = open(self._log_path, "ab")

    # ---- recovery ----------------------------------------------------------------------------
    def _load(self) -> None:
        self._id_of.clear(); self._key_of.clear(); self._meta.clear()
        self._log_bytes = 0
        if not os.path.exists(self._log_path):
            return
        good = 0
        with open(self._log_path, "rb") as f:
            for raw in f:
                if not raw.endswith(b"\n"):
                    break                                   # torn tail of a crashed append
                try:
                    rec = json.loads(raw)
                except json.JSONDecodeError:
                    break
                good += len(raw)
                if "d" in rec:
                    old = self._id_of.pop(rec["d"], None)
                    if old is not None:
                        self._key_of.pop(old, None)
                        self._meta.pop(old, None)
                else:
                    self._apply_put(int(rec["i"]), rec["k"], rec.get("m"))
        if good != os.path.getsize(self._log_path):
            with open(self._log_path, "r+b") as f:
                f.truncate(good)
        self._log_bytes = good

    def _apply_put(self, hid: int, key: str, meta: Optional[dict]) -> None:
        old = self._id_of.get(key)
        if old is not None:
            self._key_of.pop(old, None)
            self._meta.pop(old, None)
        self._id_of[key] = hid
        self._key_of[hid] = key
        if meta:
            self._meta[hid] = meta

    # ---- writes ------------------------------------------------------------------------------
    def _append_log(self, data: bytes) -> None:
        self._log.write(data)
        self._log.flush()
        if self.fsync:
            os.fsync(self._log.fileno())
        self._log_bytes += len(data)

    def put(self, hid: int, key: str, vec: np.ndarray, meta: Optional[dict] = None) -> None:
        self.put_batch([hid], [key], np.asarray(vec, dtype=np.float32).reshape(1, -1), [meta])

    def put_batch(self, ids: Sequence[int], keys: Sequence[str], vecs: np.ndarray, metas: Sequence[Optional[dict]]) -> None:
        """Vectors first (one pwrite per run of consecutive ids, then fsync), then the log records: a record is
        only ever durable after the vector it points at."""
        vecs = np.ascontiguousarray(vecs, dtype=np.float32)
        n = len(ids)
        if n == 0:
            return
        if vecs.shape != (n, self.dim):
            raise ValueError(f"vectors must be [{n}, {self.dim}] float32")
        ids = [int(i) for i in ids]
        run0 = 0
        for j in range(1, n + 1):
            if j == n or ids[j] != ids[j - 1] + 1:
                os.pwrite(self._fd, vecs[run0:j].tobytes(), ids[run0] * self.row_bytes)
                run0 = j
        if self.fsync:
            os.fsync(self._fd)
        dumps = json.dumps
        lines = []
        for hid, key, meta in zip(ids, keys, metas):
            lines.append(dumps({"i": hid, "k": key, "m": meta} if meta else {"i": hid, "k": key}, ensure_ascii=False))
            self._apply_put(hid, key, meta)
        self._append_log(("\n".join(lines) + "\n").encode("utf-8"))

    def delete(self, key: str) -> int:
        """-> the id the key had, or -1."""
        old = self._id_of.pop(key, None)
        if old is None:
            return -1
        self._key_of.pop(old, None)
        self._meta.pop(old, None)
        self._append_log((json.dumps({"d": key}, ensure_ascii=False) + "\n").encode("utf-8"))
        return old

    # ---- reads -------------------------------------------------------------------------------
    def id_of(self, key: str) -> int:
        return self._id_of.get(key, -1)

    def key_of(self, hid: int) -> str:
        return self._key_of.get(hid, "")

    def metadata(self, hid: int) -> dict:
        return self._meta.get(hid, {})

    def vector(self, hid: int) -> np.ndarray:
        raw = os.pread(self._fd, self.row_bytes, hid * self.row_bytes)
        if len(raw) != self.row_bytes:
            raise KeyError(f"no vector stored for id {hid}")
        return np.frombuffer(raw, dtype=np.float32)

    def vectors(self, ids: Iterable[int]) -> np.ndarray:
        ids = list(ids)
        out = np.empty((len(ids), self.dim), dtype=np.float32)
        for r, hid in enumerate(ids):
            out[r] = self.vector(hid)
        return out

    def __len__(self) -> int:
        return len(self._id_of)

    def __contains__(self, key: str) -> bool:
        return key in self._id_of

    def keys(self) -> List[str]:
        return list(self._id_of)

    # ---- checkpoints -------------------------------------------------------------------------
    def position(self) -> dict:
        """What a checkpoint records: every byte of the log before it is part of the checkpoint."""
        return {"log_bytes": self._log_bytes}

    def rollback(self, pos: dict) -> None:
        """Forget everything appended after `pos` (recovery: the WAL replay re-applies it)."""
        want = int(pos.get("log_bytes", 0))
        if want >= self._log_bytes:
            return
        self._log.close()
        with open(self._log_path, "r+b") as f:
            f.truncate(want)
        self._load()
        self._log = open(self._log_path, "ab")

    def flush(self) -> None:
        self._log.flush()
        os.fsync(self._log.fileno())
        os.fsync(self._fd)

    def close(self) -> None:
        try:
            self._log.close()
        finally:
            if self._fd is not None:
                os.close(self._fd)
                self._fd = None
