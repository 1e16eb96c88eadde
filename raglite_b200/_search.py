"""Drop-in ``vector_search`` / ``rerank_chunks`` over the device-resident index.

Signatures follow the reference (``raglite/_search.py:36-43`` and ``:364-366``); the arithmetic that
the reference delegates to DuckDB SQL runs in the CUDA library instead.  ``vector_search_batch`` is
the batched entry the benchmark configs use (the reference API is single-query, ``_search.py:54-56``).
"""

from __future__ import annotations

import contextlib
from collections.abc import Sequence
from dataclasses import dataclass
from typing import Any

import numpy as np
import torch

from ._config import RAGLiteConfig
from ._index import Chunk, CorpusIndex, get_index, search_to_host
from ._typing import ChunkId, FloatVector, MetadataFilter

REFERENCE_CHUNK_MAX_SIZE = 2048  # RAGLiteConfig.chunk_max_size class default (_config.py:67)


def num_hits_rule(num_results: int, oversample: int, chunk_max_size: int) -> int:
    """``_search.py:66-67``."""
    corrected_oversample = oversample * chunk_max_size / REFERENCE_CHUNK_MAX_SIZE
    return round(corrected_oversample) * max(num_results, 10)


def _adapt_metadata(metadata_filter: MetadataFilter | None) -> dict[str, list[Any]] | None:
    """Normalise filter values to lists (``_database.py:51-55``)."""
    if not metadata_filter:
        return None
    return {k: (list(v) if isinstance(v, (list, tuple)) else [v]) for k, v in metadata_filter.items()}


FILTER_FIRST_MAX_ROWS = 100_000   # metadata_count <= 100_000: filter, then rank (_search.py:105)
RANK_FIRST_LIMIT = 1_000_000      # otherwise: the 1_000_000 nearest vectors, then the filter (_search.py:126)


def _filter_on_device(index: Any, metadata_filter: dict[str, list[Any]] | None) -> tuple[torch.Tensor | None, int]:
    """The metadata filter as a per-chunk byte mask on the device plus the number of live rows it
    matches on this shard (``CorpusIndex.filter_chunks``: an inverted index over the chunk metadata,
    built once and cached per filter -- a search never walks the chunk table on the host)."""
    if not metadata_filter:
        return None, 0
    local: CorpusIndex = getattr(index, "local", index)
    return local.filter_chunks(metadata_filter)


def vector_search_batch(  # noqa: PLR0913
    queries: np.ndarray | torch.Tensor,
    *,
    num_results: int = 3,
    oversample: int = 4,
    metadata_filter: MetadataFilter | None = None,
    config: RAGLiteConfig | None = None,
    index: Any | None = None,
    exact_maxsim: bool = False,
    algo: str = "auto",
    queries_are_fp16: bool = False,
) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Batched ``vector_search``: ``queries`` is ``[B, d]`` (host or device).

    Returns host arrays ``(chunk_index[B, k] int64 (-1 padded), sim[B, k] float32, count[B])``.
    ``exact_maxsim=True`` ranks by exact per-chunk MaxSim instead of the reference's
    top-``num_hits``-vectors semantics.  Host work per call is O(B): the query upload, kernel launches,
    one pinned device->host copy of the results and one stream synchronisation.
    """
    config = config or RAGLiteConfig()
    index = index if index is not None else get_index(config)
    if index is None:
        raise ValueError(f"No index registered for db_url={config.db_url!r}; use raglite_b200.register_index")
    local: CorpusIndex = getattr(index, "local", index)
    Q = torch.as_tensor(queries)
    queries_are_fp16 = queries_are_fp16 or Q.dtype == torch.float16
    Q = Q.to(device=local.device, dtype=torch.float32, non_blocking=True).contiguous()
    if Q.ndim != 2:
        raise ValueError("queries must be [B, d]")
    k = int(num_results)
    B = int(Q.shape[0])
    sharded = hasattr(index, "group")
    if local.n_live_chunks == 0 and not sharded:
        return np.full((B, k), -1, np.int64), np.full((B, k), -np.inf, np.float32), np.zeros(B, np.int32)
    if config.vector_search_query_adapter and local.query_adapter is not None:
        Q = local.apply_adapter(Q, round_fp16=queries_are_fp16)
    num_hits = 0 if exact_maxsim else num_hits_rule(k, oversample, config.chunk_max_size)
    if not exact_maxsim and num_hits == 0:  # round(oversample * size / 2048) == 0 -> LIMIT 0
        return np.full((B, k), -1, np.int64), np.full((B, k), -np.inf, np.float32), np.zeros(B, np.int32)
    chunk_ok, n_match = _filter_on_device(index, _adapt_metadata(metadata_filter))
    metric = config.vector_search_distance_metric
    # Which metadata branch the reference would take (_search.py:96-143): many matching rows in a corpus
    # of more than 1M vectors -> only filtered hits among the 1M nearest vectors overall count.
    rank_first_limit = None
    if chunk_ok is not None and num_hits > 0:
        totals = [n_match, local.n_live_rows]
        if sharded:
            totals = index.sum_over_shards(torch.tensor(totals, dtype=torch.int64, device=local.device)).tolist()
        if totals[0] > FILTER_FIRST_MAX_ROWS and totals[1] > RANK_FIRST_LIMIT:
            rank_first_limit = RANK_FIRST_LIMIT
    return search_to_host(index, Q, k=k, num_hits=num_hits, metric=metric, algo=algo, chunk_ok=chunk_ok,
                          rank_first_limit=rank_first_limit)


def vector_search(
    query: str | FloatVector,
    *,
    num_results: int = 3,
    oversample: int = 4,
    metadata_filter: MetadataFilter | None = None,
    config: RAGLiteConfig | None = None,
) -> tuple[list[ChunkId], list[float]]:
    """Search chunks by multi-vector similarity -- drop-in for ``raglite.vector_search``
    (``_search.py:36-153``): embed / ravel the query, apply the query adapter, keep the
    ``num_hits`` nearest vectors, group by chunk with ``max(sim)``, return the best ``num_results``."""
    config = config or RAGLiteConfig()
    index = get_index(config)
    if index is None:
        raise ValueError(f"No index registered for db_url={config.db_url!r}; use raglite_b200.register_index")
    if config.self_query and isinstance(query, str):
        raise NotImplementedError("self_query needs an LLM and is outside the accelerated hot path")
    if isinstance(query, str):
        from ._embed import embed_strings

        q = embed_strings([query], config=config)[0, :]
    else:
        q = np.ravel(query)
    local: CorpusIndex = getattr(index, "local", index)
    ids, sims, counts = vector_search_batch(
        q[None, :], num_results=num_results, oversample=oversample, metadata_filter=metadata_filter,
        config=config, index=index, queries_are_fp16=(q.dtype == np.float16),
    )
    n = int(counts[0])
    owner = index if hasattr(index, "chunk_id_of") else local
    return [owner.chunk_id_of(int(c)) for c in ids[0, :n]], [float(s) for s in sims[0, :n]]


def retrieve_chunks(chunk_ids: Sequence[ChunkId], *, config: RAGLiteConfig | None = None) -> list[Chunk]:
    """``_search.py:283-299`` over the registered index's chunk table (order follows ``chunk_ids``)."""
    config = config or RAGLiteConfig()
    if not chunk_ids:
        return []
    index = get_index(config)
    local = getattr(index, "local", index) if index is not None else None
    if local is None or local.chunks is None:
        raise ValueError("The registered index holds no chunk texts")
    by_id = {c.id: c for c in local.live_chunks}
    return [by_id[cid] for cid in chunk_ids if cid in by_id]


def rerank_chunks(
    query: str, chunk_ids: list[ChunkId] | list[Chunk], *, config: RAGLiteConfig | None = None
) -> list[Chunk]:
    """Rerank chunks by cross-encoder relevance -- drop-in for ``raglite.rerank_chunks``
    (``_search.py:364-397``): same early exits, language routing and ``.rank(query=, docs=)`` contract."""
    config = config or RAGLiteConfig()
    chunks: list[Chunk] = (
        retrieve_chunks(chunk_ids, config=config)  # type: ignore[arg-type]
        if all(isinstance(c, ChunkId) for c in chunk_ids)
        else list(chunk_ids)  # type: ignore[arg-type]
    )
    if not config.reranker or not chunks:
        return chunks
    if isinstance(config.reranker, dict):
        langs: set[str] = set()
        try:  # the reference detects languages with langdetect (_search.py:381-383); optional here
            from langdetect import LangDetectException, detect  # type: ignore[import-not-found]

            with contextlib.suppress(LangDetectException):
                langs = {detect(str(chunk)) for chunk in chunks}
                langs.add(detect(query))
        except ModuleNotFoundError:
            langs = set()
        rerankers = config.reranker
        if len(langs) == 1 and (lang := next(iter(langs))) in rerankers:
            reranker = rerankers[lang]
        else:
            reranker = rerankers.get("other")
    else:
        reranker = config.reranker
    if reranker:
        results = reranker.rank(query=query, docs=[str(chunk) for chunk in chunks])
        chunks = [chunks[result.doc_id] for result in results.results]
    return chunks


# ---- the steps right after the hot path ------------------------------------------
@dataclass
class ChunkSpan:
    """A run of consecutive chunks of one document (reference ``_database.py:326-398``)."""

    chunks: list[Chunk]

    @property
    def document_id(self) -> str:
        return self.chunks[0].document_id if self.chunks else ""

    @property
    def content(self) -> str:
        """Front matter and heading of the first chunk, then all bodies (``_database.py:389-394``)."""
        if not self.chunks:
            return ""
        bodies = "".join(chunk.body for chunk in self.chunks)
        return f"{self.chunks[0].front_matter}\n\n{self.chunks[0].headings.strip()}\n\n{bodies}".strip()

    def __str__(self) -> str:
        return self.content


def retrieve_chunk_spans(
    chunk_ids: list[ChunkId] | list[Chunk], *, neighbors: tuple[int, ...] | None = (-1, 1),
    config: RAGLiteConfig | None = None,
) -> list[ChunkSpan]:
    """Group chunks (plus their ``neighbors`` in the same document) into contiguous spans, ordered by
    the summed reciprocal rank ``1 / (i + 1)`` of the chunks they contain (``_search.py:302-361``)."""
    if not chunk_ids:
        return []
    config = config or RAGLiteConfig()
    chunks: list[Chunk] = (
        retrieve_chunks(chunk_ids, config=config)  # type: ignore[arg-type]
        if all(isinstance(c, ChunkId) for c in chunk_ids) else list(chunk_ids)  # type: ignore[arg-type]
    )
    score = {chunk.id: 1 / (i + 1) for i, chunk in enumerate(chunks)}
    pool: dict[tuple[str, int], Chunk] = {(c.document_id, c.index): c for c in chunks}
    if neighbors:
        index = get_index(config)
        local = getattr(index, "local", index) if index is not None else None
        table = {(c.document_id, c.index): c for c in local.live_chunks} if local is not None else {}
        for c in chunks:
            for off in neighbors:
                nb = table.get((c.document_id, c.index + off))
                if nb is not None:
                    pool.setdefault((nb.document_id, nb.index), nb)
    spans: list[ChunkSpan] = []
    run: list[Chunk] = []
    for key in sorted(pool):
        c = pool[key]
        if run and (c.document_id != run[-1].document_id or c.index != run[-1].index + 1):
            spans.append(ChunkSpan(run))
            run = []
        run.append(c)
    if run:
        spans.append(ChunkSpan(run))
    spans.sort(key=lambda s: sum(score.get(c.id, 0.0) for c in s.chunks), reverse=True)
    return spans


def search_and_rerank_chunks(  # noqa: PLR0913
    query: str, *, num_results: int = 8, oversample: int = 4, search: Any = None,
    config: RAGLiteConfig | None = None, metadata_filter: MetadataFilter | None = None,
) -> list[Chunk]:
    """Search ``oversample * num_results`` chunks, rerank, keep ``num_results`` (``_search.py:400-413``).
    The reference defaults ``search`` to hybrid search; keyword search is out of scope here, so the
    default is ``vector_search``."""
    search = search or vector_search
    chunk_ids, _ = search(query, num_results=oversample * num_results, metadata_filter=metadata_filter, config=config)
    return rerank_chunks(query, chunk_ids, config=config)[:num_results]


def search_and_rerank_chunk_spans(  # noqa: PLR0913
    query: str, *, num_results: int = 8, oversample: int = 4, neighbors: tuple[int, ...] | None = (-1, 1),
    search: Any = None, config: RAGLiteConfig | None = None, metadata_filter: MetadataFilter | None = None,
) -> list[ChunkSpan]:
    """``search_and_rerank_chunks`` followed by span collation (``_search.py:416-433``)."""
    chunks = search_and_rerank_chunks(query, num_results=num_results, oversample=oversample, search=search,
                                      config=config, metadata_filter=metadata_filter)
    return retrieve_chunk_spans(chunks, neighbors=neighbors, config=config)
