"""NumPy restatement of ``raglite._search.vector_search``'s arithmetic (TEST INFRASTRUCTURE).

PARITY UNPINNED (see ``oracle/__init__.py``): the distance / ORDER BY / GROUP BY steps run inside
DuckDB in the reference; this module restates them from the reference's call sites.

Conventions
-----------
``E``          float32 ``[N, d]`` -- one row per ``chunk_embedding`` row, a chunk's vectors
               contiguous (reference ``_insert.py:247-251``).
``chunk_off``  int64 ``[C + 1]`` -- CSR row offsets: chunk ``c`` owns rows
               ``chunk_off[c]:chunk_off[c + 1]`` (variable count, ``_split_chunks.py:121``).
"""

from __future__ import annotations

import numpy as np

REFERENCE_CHUNK_MAX_SIZE = 2048  # RAGLiteConfig.chunk_max_size class default (_config.py:67).

METRICS = ("cosine", "dot", "l2")


def num_hits_rule(num_results: int, oversample: int = 4, chunk_max_size: int = 2048) -> int:
    """``_search.py:66-67``: ``round(oversample * chunk_max_size / 2048) * max(num_results, 10)``."""
    corrected_oversample = oversample * chunk_max_size / REFERENCE_CHUNK_MAX_SIZE
    return round(corrected_oversample) * max(num_results, 10)


def apply_query_adapter(A: np.ndarray | None, q: np.ndarray) -> np.ndarray:
    """``_search.py:58-62``: ``(Q @ query_embedding).astype(query_embedding.dtype)``."""
    q = np.ravel(q)
    if A is None:
        return q
    return (A @ q).astype(q.dtype)


def vector_distances(E: np.ndarray, q: np.ndarray, metric: str = "cosine") -> np.ndarray:
    """Per-row distance as DuckDB computes it on ``FLOAT[d]`` columns (``_typing.py:123-134``).

    cosine: ``array_cosine_distance``  = 1 - clamp(<e,q> / sqrt(|e|^2 |q|^2), -1, 1)
    dot:    ``array_negative_inner_product`` = -<e,q>
    l2:     ``array_distance``         = sqrt(sum (e - q)^2)
    All in float32 (the query is cast ``::FLOAT[d]``, ``_typing.py:133``).
    """
    E = np.asarray(E, dtype=np.float32)
    q = np.ravel(q).astype(np.float32)
    if metric == "cosine":
        dot = E @ q
        nl = np.einsum("ij,ij->i", E, E)
        nr = np.float32(q @ q)
        sim = dot / np.sqrt(nl * nr)
        sim = np.clip(sim, np.float32(-1.0), np.float32(1.0))
        return (np.float32(1.0) - sim).astype(np.float32)
    if metric == "dot":
        return (-(E @ q)).astype(np.float32)
    if metric == "l2":
        diff = E - q[None, :]
        return np.sqrt(np.einsum("ij,ij->i", diff, diff)).astype(np.float32)
    raise ValueError(f"Unsupported metric: {metric}")


def vector_distances_f64(E: np.ndarray, q: np.ndarray, metric: str = "cosine") -> np.ndarray:
    """Float64 version of :func:`vector_distances`, used to adjudicate near-ties."""
    E = np.asarray(E, dtype=np.float64)
    q = np.ravel(q).astype(np.float64)
    if metric == "cosine":
        sim = (E @ q) / np.sqrt(np.einsum("ij,ij->i", E, E) * (q @ q))
        return 1.0 - np.clip(sim, -1.0, 1.0)
    if metric == "dot":
        return -(E @ q)
    if metric == "l2":
        diff = E - q[None, :]
        return np.sqrt(np.einsum("ij,ij->i", diff, diff))
    raise ValueError(f"Unsupported metric: {metric}")


def row_to_chunk(chunk_off: np.ndarray, n_rows: int | None = None) -> np.ndarray:
    """Map each embedding row to the chunk that owns it."""
    chunk_off = np.asarray(chunk_off, dtype=np.int64)
    n_rows = int(chunk_off[-1]) if n_rows is None else n_rows
    return (np.searchsorted(chunk_off, np.arange(n_rows), side="right") - 1).astype(np.int64)


def vector_search_sql(  # noqa: PLR0913
    E: np.ndarray,
    chunk_off: np.ndarray,
    q: np.ndarray,
    *,
    num_results: int = 3,
    oversample: int = 4,
    chunk_max_size: int = 2048,
    metric: str = "cosine",
    adapter: np.ndarray | None = None,
    allowed_chunks: np.ndarray | None = None,
    f64: bool = False,
    filter_first_max: int = 100_000,
    rank_first_limit: int = 1_000_000,
    row_chunk: np.ndarray | None = None,
    f32_ties: bool = False,
) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Exact-scan restatement of ``vector_search`` (``_search.py:58-153``, no HNSW approximation).

    Steps: adapter apply -> per-row ``dist`` -> ``ORDER BY dist LIMIT num_hits`` (top vectors)
    -> ``GROUP BY chunk_id, max(sim)`` -> ``ORDER BY sim DESC LIMIT num_results``.
    ``allowed_chunks`` (bool ``[C]``) restates the filter-first metadata branch
    (``_search.py:105-121``) when at most ``filter_first_max`` rows match, and the rank-then-filter
    branch (``_search.py:122-143``: the ``rank_first_limit`` nearest rows, then the filter) otherwise.
    Ties are broken by row / chunk index (SQL leaves them unspecified).  ``row_chunk`` (int64 ``[N]``)
    names the owner chunk of every row directly instead of through ``chunk_off`` (a gathered subset of a
    larger table).  ``f32_ties`` (with ``f64``) rounds the float64 distance to the FLOAT DuckDB returns
    before ordering, so rows that tie in float32 are ordered by row index.

    Returns ``(chunk_index[int64], sim[float], hit_rows[int64])`` where ``hit_rows`` are the
    ``num_hits`` selected vector rows in ascending-distance order.
    """
    E = np.asarray(E)
    if E.shape[0] == 0:
        return np.zeros(0, np.int64), np.zeros(0, np.float32), np.zeros(0, np.int64)
    q = apply_query_adapter(adapter, q)
    num_hits = num_hits_rule(num_results, oversample, chunk_max_size)
    dist = (vector_distances_f64 if f64 else vector_distances)(E, q, metric)
    if f64 and f32_ties:
        dist = float_distance_of_f64(dist, metric)
    r2c = row_to_chunk(chunk_off, E.shape[0]) if row_chunk is None else np.asarray(row_chunk, dtype=np.int64)
    rows = np.arange(E.shape[0])
    if allowed_chunks is not None:
        row_ok = np.asarray(allowed_chunks, dtype=bool)[r2c]
        if int(row_ok.sum()) <= filter_first_max:      # metadata_count <= 100_000 (_search.py:105)
            rows = rows[row_ok]
        else:                                          # ORDER BY dist LIMIT 1_000_000, then the filter
            nearest = np.argsort(dist, kind="stable")[:rank_first_limit]
            rows = np.sort(nearest[row_ok[nearest]])
    order = rows[np.argsort(dist[rows], kind="stable")][:num_hits]
    ids, sims = group_hits(dist[order], r2c[order], num_results)
    return ids, sims, order.astype(np.int64)


def float_distance_of_f64(dist64: np.ndarray, metric: str) -> np.ndarray:
    """The FLOAT a float64-accurate distance becomes when returned as DuckDB's ``FLOAT`` column:
    cosine ``1 - s`` is formed in float32 from the float32-rounded similarity."""
    if metric == "cosine":
        return (np.float32(1.0) - (1.0 - dist64).astype(np.float32)).astype(np.float32)
    return dist64.astype(np.float32)


def group_hits(hit_dist: np.ndarray, hit_chunk: np.ndarray, num_results: int) -> tuple[np.ndarray, np.ndarray]:
    """``GROUP BY chunk_id -> max(sim) -> ORDER BY sim DESC LIMIT num_results`` (``_search.py:143-150``) over
    the selected vectors, given in ascending-distance order: the first occurrence of a chunk carries
    its max; chunk ties are broken by chunk index."""
    one = np.float32(1.0) if hit_dist.dtype == np.float32 else 1.0
    sim = one - hit_dist
    uniq, first = np.unique(hit_chunk, return_index=True)
    grouped_sim = sim[first]
    rank = np.lexsort((uniq, -grouped_sim.astype(np.float64)))[:num_results]
    return uniq[rank].astype(np.int64), grouped_sim[rank]


def topn_rows_blocked(blocks, Q: np.ndarray, n_keep: int, metric: str = "cosine", *, f32_ties: bool = False,
                      row_ok=None) -> list[tuple[np.ndarray, np.ndarray]]:
    """``ORDER BY dist LIMIT n_keep`` (``_search.py:75-79``) for a batch of queries over a table that is
    handed over block by block: ``blocks`` yields ``(first_row, E_block float32 [n, d])`` in row order
    (e.g. slices copied back from the device), so a corpus larger than host memory can be checked.
    Distances are float64 (``vector_distances_f64`` semantics; with ``f32_ties`` rounded to the FLOAT
    DuckDB returns before ordering); ties are broken by row index exactly as the unblocked restatement
    does.  ``row_ok(first_row, n) -> bool[n]`` optionally restricts the rows (metadata filter).
    Returns, per query, ``(rows int64, dist)`` ascending."""
    Q64 = np.asarray(Q, dtype=np.float64)
    B = Q64.shape[0]
    qq = np.einsum("ij,ij->i", Q64, Q64)
    keep_d: list[np.ndarray] = [np.zeros(0, np.float32 if f32_ties else np.float64) for _ in range(B)]
    keep_r: list[np.ndarray] = [np.zeros(0, np.int64) for _ in range(B)]
    for row0, Eb in blocks:
        E64 = np.asarray(Eb, dtype=np.float64)
        n = E64.shape[0]
        if n == 0:
            continue
        G = E64 @ Q64.T                                   # [n, B]
        if metric == "cosine":
            ee = np.einsum("ij,ij->i", E64, E64)
            D = 1.0 - np.clip(G / np.sqrt(ee[:, None] * qq[None, :]), -1.0, 1.0)
        elif metric == "dot":
            D = -G
        elif metric == "l2":
            ee = np.einsum("ij,ij->i", E64, E64)
            D = np.sqrt(np.maximum(ee[:, None] + qq[None, :] - 2.0 * G, 0.0))
            # the unblocked restatement sums (e - q)^2; recompute the few kept rows that way below
        else:
            raise ValueError(f"Unsupported metric: {metric}")
        ok = None if row_ok is None else np.asarray(row_ok(row0, n), dtype=bool)
        for b in range(B):
            d = D[:, b]
            idx = np.arange(n) if ok is None else np.nonzero(ok)[0]
            d = d[idx]
            if len(d) > n_keep:
                v = np.partition(d, n_keep - 1)[n_keep - 1]
                sel = d <= v
                d, idx = d[sel], idx[sel]
            if metric == "l2" and len(idx):
                diff = E64[idx] - Q64[b][None, :]
                d = np.sqrt(np.einsum("ij,ij->i", diff, diff))
            if f32_ties:
                d = float_distance_of_f64(d, metric)
            cd = np.concatenate([keep_d[b], d])
            cr = np.concatenate([keep_r[b], idx.astype(np.int64) + int(row0)])
            if len(cd) > n_keep:
                v = np.partition(cd, n_keep - 1)[n_keep - 1]
                sel = cd <= v
                cd, cr = cd[sel], cr[sel]
            keep_d[b], keep_r[b] = cd, cr
    out = []
    for b in range(B):
        o = np.lexsort((keep_r[b], keep_d[b]))[:n_keep]
        out.append((keep_r[b][o], keep_d[b][o]))
    return out


def maxsim_scores(
    E: np.ndarray, chunk_off: np.ndarray, q: np.ndarray, metric: str = "cosine", *, f64: bool = True
) -> np.ndarray:
    """Per-chunk MaxSim score ``max_j sim(e_j, q)`` -- the NumPy spelling the reference uses in the
    adapter fit (``_query_adapter.py:172-183``: ``argmax(chunk.embedding_matrix @ q)``), extended to
    every metric through ``sim = 1 - dist`` (``_search.py:72``)."""
    dist = (vector_distances_f64 if f64 else vector_distances)(E, q, metric)
    sim = (1.0 if f64 else np.float32(1.0)) - dist
    chunk_off = np.asarray(chunk_off, dtype=np.int64)
    return np.maximum.reduceat(sim, chunk_off[:-1])


def maxsim_topk_exact(
    E: np.ndarray, chunk_off: np.ndarray, q: np.ndarray, k: int, metric: str = "cosine"
) -> tuple[np.ndarray, np.ndarray]:
    """Exact per-chunk MaxSim ranking, float64 accumulation, ties by chunk index."""
    if np.asarray(E).shape[0] == 0:
        return np.zeros(0, np.int64), np.zeros(0, np.float64)
    s = maxsim_scores(E, chunk_off, q, metric, f64=True)
    rank = np.lexsort((np.arange(len(s)), -s))[:k]
    return rank.astype(np.int64), s[rank]


def vector_search_batch(  # noqa: PLR0913
    E: np.ndarray,
    chunk_off: np.ndarray,
    Q: np.ndarray,
    *,
    num_results: int,
    oversample: int = 4,
    chunk_max_size: int = 2048,
    metric: str = "cosine",
    adapter: np.ndarray | None = None,
    exact_maxsim: bool = False,
) -> list[tuple[np.ndarray, np.ndarray]]:
    """Loop the single-query reference semantics over a batch (the reference API is single-query)."""
    out = []
    for q in np.asarray(Q):
        if exact_maxsim:
            out.append(maxsim_topk_exact(E, chunk_off, apply_query_adapter(adapter, q), num_results, metric))
        else:
            ids, sims, _ = vector_search_sql(
                E, chunk_off, q, num_results=num_results, oversample=oversample,
                chunk_max_size=chunk_max_size, metric=metric, adapter=adapter,
            )
            out.append((ids, sims))
    return out


def blas_batch_topk(  # noqa: PLR0913
    E: np.ndarray, vecs_per_chunk: int, Q: np.ndarray, k: int, *, num_hits: int = 0, metric: str = "cosine"
) -> tuple[np.ndarray, np.ndarray]:
    """Batched best case of the reference's arithmetic for the CPU baseline: one sgemm
    ``S = Q E^T`` on all host cores, cosine scaling, then either the SQL semantics
    (top-``num_hits`` vectors -> group max) or plain reshape-max, then ``argpartition`` top-k.
    Fixed ``vecs_per_chunk`` only (the benchmark shapes)."""
    E = np.asarray(E, dtype=np.float32)
    Q = np.asarray(Q, dtype=np.float32)
    S = Q @ E.T
    if metric == "cosine":
        S /= np.sqrt(np.einsum("ij,ij->i", E, E))[None, :]
        S /= np.sqrt(np.einsum("ij,ij->i", Q, Q))[:, None]
    B, N = S.shape
    C = N // vecs_per_chunk
    if num_hits:
        nh = min(num_hits, N)
        part = np.argpartition(-S, nh - 1, axis=1)[:, :nh]
        ids = np.full((B, k), -1, np.int64)
        sc = np.full((B, k), -np.inf, np.float32)
        for b in range(B):
            rows = part[b][np.argsort(-S[b, part[b]], kind="stable")]
            ch = rows // vecs_per_chunk
            uniq, first = np.unique(ch, return_index=True)
            sims = S[b, rows[first]]
            o = np.lexsort((uniq, -sims))[:k]
            ids[b, : len(o)] = uniq[o]
            sc[b, : len(o)] = sims[o]
        return ids, sc
    M = S[:, : C * vecs_per_chunk].reshape(B, C, vecs_per_chunk).max(axis=2)
    kk = min(k, C)
    part = np.argpartition(-M, kk - 1, axis=1)[:, :kk]
    ps = np.take_along_axis(M, part, axis=1)
    o = np.argsort(-ps, axis=1, kind="stable")
    return np.take_along_axis(part, o, axis=1), np.take_along_axis(ps, o, axis=1)
