"""B200-native exact-kNN shard search: the datanode search hot path of
f1ybaozii/Distributed-Vector-Database rebuilt on sm_100a CUDA behind a C ABI.

Importable as ``dvdb_b200`` (see the alias module at the repo root); the directory keeps the
name the build contract asks for.  Importing the package loads no native code; the first use of
``Index`` / ``merge_topk`` loads ``libvdb_b200.so`` and fails loudly when it (or a GPU) is missing."""
from . import _ffi, kvstore
from .coordinator import LocalCoordinator, PeerExchange, ShardedIndex, ShardedSearcher, merge_search_results
from .handler import GpuVectorNodeHandler
from .index import Index, launch_count, merge_topk, pinned_empty
from .sharding import assign_shards_to_nodes, get_shard_id
from .ttypes import Response, SearchRequest, SearchResult, VectorData
from .wal import WALManager

__all__ = ["Index", "merge_topk", "pinned_empty", "launch_count", "_ffi", "GpuVectorNodeHandler", "LocalCoordinator",
           "ShardedSearcher", "ShardedIndex", "PeerExchange", "merge_search_results", "WALManager", "get_shard_id", "assign_shards_to_nodes",
           "VectorData", "SearchRequest", "SearchResult", "Response"]
