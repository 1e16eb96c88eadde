"""Seeded synthetic corpora / queries shared by tests, smoke and bench."""

from __future__ import annotations

import numpy as np


def make_corpus(n_chunks: int, vecs, dim: int, seed: int = 0, *, fp16_round: bool = False, normalize: bool = True):
    """Corpus ``E [N, dim]`` float32 + CSR ``chunk_off [C+1]``.  ``vecs`` is an int (fixed vectors
    per chunk) or ``(lo, hi)`` for variable counts U{lo..hi} (``_split_chunks.py:121``)."""
    rng = np.random.default_rng(seed)
    if isinstance(vecs, int):
        counts = np.full(n_chunks, vecs, dtype=np.int64)
    else:
        counts = rng.integers(vecs[0], vecs[1] + 1, size=n_chunks).astype(np.int64)
    chunk_off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    E = rng.standard_normal((int(chunk_off[-1]), dim)).astype(np.float32)
    if normalize:
        E /= np.linalg.norm(E, axis=1, keepdims=True)
    if fp16_round:  # mirrors the reference storing fp16-rounded values (_embed.py:140)
        E = E.astype(np.float16).astype(np.float32)
    return E, chunk_off


def make_queries(E: np.ndarray, n_queries: int, seed: int = 1, noise: float = 0.3, frac_random: float = 0.25):
    """Queries near random corpus rows (guaranteed neighbour) mixed with pure-random ones."""
    rng = np.random.default_rng(seed)
    d = E.shape[1]
    Q = rng.standard_normal((n_queries, d)).astype(np.float32)
    Q /= np.linalg.norm(Q, axis=1, keepdims=True)
    n_near = n_queries - int(round(frac_random * n_queries))
    if E.shape[0] and n_near:
        rows = rng.integers(0, E.shape[0], size=n_near)
        Q[:n_near] = E[rows] + noise * Q[:n_near]
        Q[:n_near] /= np.linalg.norm(Q[:n_near], axis=1, keepdims=True)
    return Q.astype(np.float32)


def random_orthogonal(d: int, seed: int = 2) -> np.ndarray:
    """Orthogonal adapter ``U @ VT`` as the cosine fit produces (``_query_adapter.py:204-205``)."""
    rng = np.random.default_rng(seed)
    U, _, VT = np.linalg.svd(rng.standard_normal((d, d)), full_matrices=False)
    return U @ VT
