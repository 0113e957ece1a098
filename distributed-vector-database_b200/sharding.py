"""Key -> shard -> node routing, as the coordinator does it (reference
src/utils/shared_utils.py:4-21, Config/storage_config.py:3)."""
from __future__ import annotations

import hashlib
from typing import Dict, List, Sequence

SHARD_COUNT = 4          # Config/storage_config.py:3
REPLICA_COUNT = 2        # Config/storage_config.py:4
VECTOR_DIM = 512         # Config/storage_config.py:2


def get_shard_id(key: str, shard_count: int = SHARD_COUNT) -> int:
    """md5(key) as an integer modulo the shard count (shared_utils.py:4-7)."""
    return int(hashlib.md5(key.encode()).hexdigest(), 16) % shard_count


def assign_shards_to_nodes(nodes: Sequence, shard_count: int = SHARD_COUNT, replica_count: int = REPLICA_COUNT) -> Dict[int, dict]:
    """Round-robin masters, the next `replica_count` nodes as slaves (shared_utils.py:9-21)."""
    mapping: Dict[int, dict] = {}
    if not nodes:
        return mapping
    for shard_id in range(shard_count):
        mapping[shard_id] = {
            "master": nodes[shard_id % len(nodes)],
            "slaves": [nodes[(shard_id + i) % len(nodes)] for i in range(1, replica_count + 1)],
        }
    return mapping


def rank_of_key(key: str, world_size: int) -> int:
    """GPU g == datanode g with SHARD_COUNT = world size: the rank that owns `key`."""
    return get_shard_id(key, world_size) % world_size


def contiguous_range(n_rows: int, rank: int, world_size: int):
    """Row range of `rank` when a synthetic set is split by contiguous blocks (bench only)."""
    return n_rows * rank // world_size, n_rows * (rank + 1) // world_size
