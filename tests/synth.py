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


# ---- clustered / anisotropic corpora (what real embedding tables look like, unlike iid Gaussians) ---------
CLUSTER_SPREADS = (0.02, 0.05, 0.1, 0.3)   # |noise| relative to the unit-norm cluster centre


def clustered_plan(n_rows: int, seed: int, *, mean_cluster: int = 512, max_cluster: int = 8192, background: float = 0.25):
    """Row -> cluster assignment shared by the NumPy and torch generators: cluster sizes are geometric
    (mean ``mean_cluster`` rows, capped at ``max_cluster``), a ``background`` fraction of the rows is not
    clustered (-1).  Rows of a cluster are contiguous runs scattered by a seeded permutation of runs, so a
    chunk's vectors mostly share a cluster (as sentences of one passage do)."""
    rng = np.random.default_rng(seed)
    n_bg = int(round(background * n_rows))
    n_cl = n_rows - n_bg
    sizes = []
    left = n_cl
    while left > 0:
        s = int(min(max_cluster, max(1, rng.geometric(1.0 / mean_cluster)), left))
        sizes.append(s)
        left -= s
    sizes = np.asarray(sizes + [n_bg], dtype=np.int64)           # last run = the background rows
    labels = np.concatenate([np.arange(len(sizes) - 1), [-1]])
    order = rng.permutation(len(sizes))
    cluster_of_row = np.repeat(labels[order], sizes[order])
    return cluster_of_row.astype(np.int64), len(sizes) - 1


def make_clustered_corpus(n_chunks: int, vecs, dim: int, seed: int = 0, *, mean_cluster: int = 512,
                          max_cluster: int = 8192, background: float = 0.25, rank: int = 16, fp16_round: bool = True):
    """Mixture of tight clusters (thousands of near-duplicates: spreads ``CLUSTER_SPREADS``) plus a
    low-rank anisotropic background, unit norm, rounded through float16 like RAGLite's stored embeddings
    (``_embed.py:140``).  Returns ``(E float32 [N, dim], chunk_off, cluster_of_row)``."""
    rng = np.random.default_rng(seed)
    if isinstance(vecs, int):
        counts = np.full(n_chunks, vecs, dtype=np.int64)
    else:
        counts = rng.integers(vecs[0], vecs[1] + 1, size=n_chunks).astype(np.int64)
    chunk_off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    N = int(chunk_off[-1])
    cl, n_clusters = clustered_plan(N, seed + 1, mean_cluster=mean_cluster, max_cluster=max_cluster, background=background)
    centers = rng.standard_normal((max(n_clusters, 1), dim)).astype(np.float32)
    centers /= np.linalg.norm(centers, axis=1, keepdims=True)
    spread = np.asarray(CLUSTER_SPREADS, np.float32)[np.arange(max(n_clusters, 1)) % len(CLUSTER_SPREADS)]
    basis = rng.standard_normal((rank, dim)).astype(np.float32) / np.sqrt(dim)
    noise = rng.standard_normal((N, dim)).astype(np.float32) / np.float32(np.sqrt(dim))
    E = np.empty((N, dim), np.float32)
    is_bg = cl < 0
    E[~is_bg] = centers[cl[~is_bg]] + spread[cl[~is_bg]][:, None] * noise[~is_bg]
    z = rng.standard_normal((int(is_bg.sum()), rank)).astype(np.float32)
    E[is_bg] = z @ basis + 0.2 * noise[is_bg]
    E /= np.linalg.norm(E, axis=1, keepdims=True)
    if fp16_round:
        E = E.astype(np.float16).astype(np.float32)
    return E, chunk_off, cl
