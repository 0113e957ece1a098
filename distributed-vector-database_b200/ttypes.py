"""Wire types of vector_db.thrift, field for field (reference src/vector_db.thrift:13-49,
generated code src/vector_db/ttypes.py:19-503).  The Thrift runtime is not a dependency of the
search path: when the generated module is importable the handlers accept its instances as well
(everything is duck-typed on the attribute names below)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional


@dataclass
class VectorData:                      # vector_db.thrift:13-18
    key: Optional[str] = None
    vector: Optional[List[float]] = None
    metadata: Optional[Dict[str, str]] = None
    timestamp: int = 0


@dataclass
class SearchRequest:                   # vector_db.thrift:23-28
    query_vector: Optional[List[float]] = None
    top_k: int = 5
    filter: Optional[Dict[str, str]] = None
    threshold: float = 0.0


@dataclass
class SearchResult:                    # vector_db.thrift:33-39
    keys: Optional[List[str]] = None
    scores: Optional[List[float]] = None
    vectors: Optional[List[VectorData]] = None
    metadatas: Optional[List[Dict[str, str]]] = None


@dataclass
class Response:                        # vector_db.thrift:44-49
    success: Optional[bool] = None
    message: str = ""
    vector_data: Optional[VectorData] = None
    search_result: Optional[SearchResult] = None
