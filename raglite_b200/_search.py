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
from ._index import Chunk, CorpusIndex, PendingSearch, get_index, search_async, search_to_host
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


def _plan_search(  # noqa: PLR0913
    queries: Any, *, num_results: int, oversample: int, metadata_filter: MetadataFilter | None, config: RAGLiteConfig | None,
    index: Any | None, exact_maxsim: bool, queries_are_fp16: bool,
) -> tuple[Any, torch.Tensor, tuple | None, dict[str, Any], Any]:
    """Argument handling shared by the synchronous and the asynchronous batched search: returns
    ``(index, Q (as given, not yet on the device), empty result or None, search kwargs, prepare(Q_device))``."""
    config = config or RAGLiteConfig()
    index = index if index is not None else get_index(config)
    if index is None:
        raise ValueError(f"No index registered for db_url={config.db_url!r}; use raglite_b200.register_index")
    local: CorpusIndex = getattr(index, "local", index)
    Q = torch.as_tensor(queries)
    queries_are_fp16 = queries_are_fp16 or Q.dtype == torch.float16
    if Q.ndim != 2:
        raise ValueError("queries must be [B, d]")
    k = int(num_results)
    B = int(Q.shape[0])
    sharded = hasattr(index, "group")
    empty = (np.full((B, k), -1, np.int64), np.full((B, k), -np.inf, np.float32), np.zeros(B, np.int32))
    if local.n_live_chunks == 0 and not sharded:
        return index, Q, empty, {}, None
    num_hits = 0 if exact_maxsim else num_hits_rule(k, oversample, config.chunk_max_size)
    if not exact_maxsim and num_hits == 0:  # round(oversample * size / 2048) == 0 -> LIMIT 0
        return index, Q, empty, {}, None
    prepare = None
    if config.vector_search_query_adapter and local.query_adapter is not None:
        def prepare(Qd: torch.Tensor) -> torch.Tensor:  # (A @ q).astype(q.dtype), _search.py:58-62
            return local.apply_adapter(Qd, round_fp16=queries_are_fp16)
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
    return index, Q, None, dict(k=k, num_hits=num_hits, metric=metric, chunk_ok=chunk_ok, rank_first_limit=rank_first_limit), prepare


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
    index, Q, empty, kw, prepare = _plan_search(queries, num_results=num_results, oversample=oversample,
                                                metadata_filter=metadata_filter, config=config, index=index,
                                                exact_maxsim=exact_maxsim, queries_are_fp16=queries_are_fp16)
    if empty is not None:
        return empty
    local: CorpusIndex = getattr(index, "local", index)
    Q = Q.to(device=local.device, dtype=torch.float32, non_blocking=True).contiguous()
    if prepare is not None:
        Q = prepare(Q)
    return search_to_host(index, Q, algo=algo, **kw)


def vector_search_batch_async(  # noqa: PLR0913
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
) -> PendingSearch:
    """``vector_search_batch`` that returns at once with a :class:`PendingSearch`; ``.result()`` gives the same three
    host arrays.  Each call runs on a stream of its own (query upload, kernels, result download into a private
    pinned buffer), so a caller that keeps two or three batches in flight -- what a retrieval server does -- never
    leaves the GPU idle between batches: the host-side work of batch i + 1 and the copies of batch i hide under the
    corpus scan, and the small kernels at either end of a search fill the scan's tail.  (The reference reaches the same
    concurrency with its thread pool, _rag.py:317; here one thread suffices.)"""
    index, Q, empty, kw, prepare = _plan_search(queries, num_results=num_results, oversample=oversample,
                                                metadata_filter=metadata_filter, config=config, index=index,
                                                exact_maxsim=exact_maxsim, queries_are_fp16=queries_are_fp16)
    if empty is not None:
        return PendingSearch(index, Q, {}, None, None, int(Q.shape[0]), int(num_results), ready=empty)
    return search_async(index, Q, algo=algo, prepare=prepare, **kw)


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


# ---- fusion and span collation batched on device chunk indices -----------------------------------------------
_KEYWORD_SEARCH: dict[str, Any] = {}


def register_keyword_search(config_or_url: Any, fn: Any) -> None:
    """Provide the BM25 keyword search for a database (the reference runs it as SQL full-text search inside
    DuckDB / PostgreSQL, ``_search.py:156-230`` -- storage-engine territory, out of scope here): any callable with
    ``keyword_search``'s signature ``(query, *, num_results, metadata_filter, config) -> (chunk_ids, scores)``."""
    _KEYWORD_SEARCH[str(getattr(config_or_url, "db_url", config_or_url))] = fn


def rrf_fuse_device(rankings: torch.Tensor, weights: Sequence[float], *, k: float = 60.0, num_results: int | None = None
                    ) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """``rl_rrf_fuse``: Reciprocal Rank Fusion (``_search.py:233-254``) of ``rankings`` -- int64 ``[B, R, L]`` chunk
    indices on the device, ``-1`` padded -- for a whole batch in one launch.  Returns device tensors
    ``(ids [B, K], score float64 [B, K], count [B])``, best first, ties in first-appearance order."""
    from . import _lib

    if rankings.dtype != torch.int64 or rankings.ndim != 3 or not rankings.is_cuda:
        raise ValueError("rankings must be a CUDA int64 [B, R, L] tensor")
    B, R, L = (int(x) for x in rankings.shape)
    if len(weights) != R:
        raise ValueError("The number of weights must match the number of rankings.")
    K = int(num_results) if num_results is not None else R * L
    dev = rankings.device
    w = torch.tensor(list(weights), dtype=torch.float64, device=dev)
    out_ids = torch.empty((B, K), dtype=torch.int64, device=dev)
    out_score = torch.empty((B, K), dtype=torch.float64, device=dev)
    out_count = torch.empty((B,), dtype=torch.int32, device=dev)
    lib = _lib.load()
    with torch.cuda.device(dev):
        _lib.check(lib.rl_rrf_fuse(rankings.contiguous().data_ptr(), w.data_ptr(), B, R, L, float(k), K, out_ids.data_ptr(),
                                   out_score.data_ptr(), out_count.data_ptr(), torch.cuda.current_stream().cuda_stream),
                   "rl_rrf_fuse")
    return out_ids, out_score, out_count


def reciprocal_rank_fusion(rankings: Sequence[Sequence[ChunkId]], *, k: int = 60, weights: Sequence[float] | None = None
                           ) -> tuple[list[ChunkId], list[float]]:
    """Drop-in ``reciprocal_rank_fusion`` (``_search.py:233-254``): the ids are interned to integers, fused by
    ``rl_rrf_fuse`` on the device (float64, the reference's summation order) and mapped back."""
    weights = [1.0] * len(rankings) if weights is None else list(weights)
    if len(weights) != len(rankings):
        raise ValueError("The number of weights must match the number of rankings.")
    L = max((len(r) for r in rankings), default=0)
    if L == 0:
        return [], []
    intern: dict[ChunkId, int] = {}
    names: list[ChunkId] = []
    table = np.full((1, len(rankings), L), -1, dtype=np.int64)
    for r, ranking in enumerate(rankings):
        for i, cid in enumerate(ranking):
            j = intern.get(cid)
            if j is None:
                j = intern[cid] = len(names)
                names.append(cid)
            table[0, r, i] = j
    ids, score, count = rrf_fuse_device(torch.from_numpy(table).cuda(), weights, k=float(k))
    n = int(count[0])
    return [names[int(j)] for j in ids[0, :n].tolist()], [float(x) for x in score[0, :n].tolist()]


def hybrid_search(  # noqa: PLR0913
    query: str, *, num_results: int = 3, oversample: int = 2, vector_search_weight: float = 0.75,
    keyword_search_weight: float = 0.25, metadata_filter: MetadataFilter | None = None, config: RAGLiteConfig | None = None,
) -> tuple[list[ChunkId], list[float]]:
    """Drop-in ``hybrid_search`` (``_search.py:257-280``): vector search on the device index, the registered keyword
    search (``register_keyword_search``), Reciprocal Rank Fusion of the two rankings on the device."""
    config = config or RAGLiteConfig()
    keyword_search = _KEYWORD_SEARCH.get(str(config.db_url))
    if keyword_search is None:
        raise ValueError("hybrid_search needs a keyword search for this database: raglite_b200.register_keyword_search")
    vs_ids, _ = vector_search(query, num_results=oversample * num_results, metadata_filter=metadata_filter, config=config)
    ks_ids, _ = keyword_search(query, num_results=oversample * num_results, metadata_filter=metadata_filter, config=config)
    ids, score = reciprocal_rank_fusion([vs_ids, list(ks_ids)], weights=[vector_search_weight, keyword_search_weight])
    return ids[:num_results], score[:num_results]


def collate_spans_device(index: Any, ranked: torch.Tensor, *, neighbors: tuple[int, ...] | None = (-1, 1)
                         ) -> dict[str, torch.Tensor]:
    """``rl_span_collate`` for a batch of ranked chunk-index lists (int64 ``[B, M]`` on the device, ``-1`` padded,
    LOCAL chunk indices of the index): the neighbour join, dedup, run cutting and span ranking of
    ``retrieve_chunk_spans`` (``_search.py:323-360``) in one launch.  Returns device tensors ``member [B, cap]``,
    ``span_start`` / ``span_len`` / ``span_score [B, cap]``, ``n_span [B]``, ``n_member [B]``."""
    from . import _lib

    local: CorpusIndex = getattr(index, "local", index)
    tabs = local.span_tables()
    B, M = (int(x) for x in ranked.shape)
    nb = torch.tensor(list(neighbors or ()), dtype=torch.int32, device=local.device)
    cap = M * (1 + int(nb.numel()))
    dev = local.device
    out = {"member": torch.empty((B, cap), dtype=torch.int64, device=dev),
           "span_start": torch.empty((B, cap), dtype=torch.int32, device=dev),
           "span_len": torch.empty((B, cap), dtype=torch.int32, device=dev),
           "span_score": torch.empty((B, cap), dtype=torch.float64, device=dev),
           "n_span": torch.empty((B,), dtype=torch.int32, device=dev), "n_member": torch.empty((B,), dtype=torch.int32, device=dev)}
    lib = _lib.load()
    with torch.cuda.device(dev):
        _lib.check(lib.rl_span_collate(
            ranked.contiguous().data_ptr(), B, M, tabs["chunk_doc"].data_ptr(), tabs["chunk_pos"].data_ptr(),
            tabs["chunk_alive"].data_ptr(), tabs["sorted_key"].data_ptr(), tabs["sorted_chunk"].data_ptr(),
            int(tabs["sorted_chunk"].numel()), nb.data_ptr() if nb.numel() else None, int(nb.numel()), out["member"].data_ptr(),
            out["span_start"].data_ptr(), out["span_len"].data_ptr(), out["span_score"].data_ptr(), out["n_span"].data_ptr(),
            out["n_member"].data_ptr(), torch.cuda.current_stream().cuda_stream), "rl_span_collate")
    return out


def retrieve_chunk_spans(
    chunk_ids: list[ChunkId] | list[Chunk], *, neighbors: tuple[int, ...] | None = (-1, 1),
    config: RAGLiteConfig | None = None,
) -> list[ChunkSpan]:
    """Group chunks (plus their ``neighbors`` in the same document) into contiguous spans, ordered by the summed
    reciprocal rank ``1 / (i + 1)`` of the chunks they contain (``_search.py:302-361``).  With a registered index
    that holds the ``Chunk`` records the whole collation runs in ``rl_span_collate`` on chunk indices; the
    host only maps ids to indices and indices back to records."""
    if not chunk_ids:
        return []
    config = config or RAGLiteConfig()
    index = get_index(config)
    local = getattr(index, "local", index) if index is not None else None
    if local is not None and local.chunks is not None and local.chunk_ids is not None:
        ids = [c if isinstance(c, ChunkId) else c.id for c in chunk_ids]
        pos = local._positions()
        idx = [pos[c] for c in ids if c in pos and local._chunk_alive[pos[c]]]
        if len(idx) == len(ids):
            ranked = torch.tensor([idx], dtype=torch.int64, device=local.device)
            out = collate_spans_device(local, ranked, neighbors=neighbors)
            n_span = int(out["n_span"][0])
            member = out["member"][0].tolist()
            start, length = out["span_start"][0, :n_span].tolist(), out["span_len"][0, :n_span].tolist()
            return [ChunkSpan([local.chunks[member[s + j]] for j in range(n)]) for s, n in zip(start, length, strict=True)]
    return _retrieve_chunk_spans_host(chunk_ids, neighbors=neighbors, config=config)


def _retrieve_chunk_spans_host(
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
