"""raglite_b200 -- B200-native (sm_100a) implementation of RAGLite's retrieval hot path.

Drop-in surface (reference ``raglite/__init__.py`` names for this path): ``RAGLiteConfig``,
``vector_search``, ``rerank_chunks``, ``embed_strings``; plus the device-resident ``CorpusIndex`` /
``ShardedIndex`` that replace the database for this path and the batched ``vector_search_batch``.
"""

from ._config import RAGLiteConfig
from ._embed import embed_strings, register_token_embedder
from ._index import Chunk, CorpusIndex, get_index, merge_hits, register_index, unregister_index
from ._query_adapter import update_query_adapter
from ._search import (
    ChunkSpan,
    collate_spans_device,
    hybrid_search,
    reciprocal_rank_fusion,
    register_keyword_search,
    rerank_chunks,
    retrieve_chunk_spans,
    retrieve_chunks,
    rrf_fuse_device,
    search_and_rerank_chunk_spans,
    search_and_rerank_chunks,
    vector_search,
    vector_search_batch,
    vector_search_batch_async,
)

__all__ = [
    "Chunk",
    "ChunkSpan",
    "CorpusIndex",
    "RAGLiteConfig",
    "collate_spans_device",
    "hybrid_search",
    "register_keyword_search",
    "rrf_fuse_device",
    "embed_strings",
    "get_index",
    "merge_hits",
    "register_index",
    "reciprocal_rank_fusion",
    "register_token_embedder",
    "rerank_chunks",
    "retrieve_chunk_spans",
    "retrieve_chunks",
    "search_and_rerank_chunk_spans",
    "search_and_rerank_chunks",
    "unregister_index",
    "update_query_adapter",
    "vector_search",
    "vector_search_batch",
    "vector_search_batch_async",
]
