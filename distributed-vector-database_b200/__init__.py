"""B200-native exact-kNN shard search: the datanode search hot path of
f1ybaozii/Distributed-Vector-Database rebuilt on sm_100a CUDA behind a C ABI.

Importable as ``dvdb_b200`` (see the alias module at the repo root); the directory keeps the
name the build contract asks for."""
from . import _ffi
from .index import Index, launch_count, merge_topk

__all__ = ["Index", "merge_topk", "launch_count", "_ffi"]
