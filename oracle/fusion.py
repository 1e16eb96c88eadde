"""Restatement of the steps right after the scan (TEST INFRASTRUCTURE): Reciprocal Rank Fusion and the ranking
half of ``retrieve_chunk_spans`` (reference ``_search.py:233-254`` and ``:323-360``).

``reciprocal_rank_fusion`` is PINNED: ``tools/make_golden_from_reference.py`` executes the reference's own function
(its source text, extracted from ``/root/reference/src/raglite/_search.py``) on seeded rankings and stores inputs
and outputs in ``tests/golden/rrf.npz``; ``tests/test_fusion.py`` checks this restatement against them.
``collate_chunk_spans`` restates the list manipulation after the reference's database lookups (unpinned: the
reference function cannot run without its SQL session; the logic restated here is the pure-Python tail)."""

from __future__ import annotations

from collections import defaultdict
from itertools import groupby


def reciprocal_rank_fusion(rankings, *, k: int = 60, weights=None):
    """``_search.py:233-254``."""
    if weights is None:
        weights = [1.0] * len(rankings)
    if len(weights) != len(rankings):
        raise ValueError("The number of weights must match the number of rankings.")
    score = defaultdict(float)
    for ranking, weight in zip(rankings, weights, strict=True):
        for i, cid in enumerate(ranking):
            score[cid] += weight / (k + i)
    if not score:
        return [], []
    ids, sc = zip(*sorted(score.items(), key=lambda x: x[1], reverse=True), strict=True)
    return list(ids), list(sc)


def collate_chunk_spans(retrieved, table, neighbors=(-1, 1)):
    """``_search.py:323-360`` on plain tuples.  ``retrieved``: chunk keys ``(document_id, index)`` in relevance order;
    ``table``: the set (or dict) of all existing ``(document_id, index)`` keys (what the neighbour query can find).
    Returns ``[(members, score)]``: spans as lists of keys, ordered by descending aggregate relevance (stable)."""
    chunks = list(retrieved)
    score = {c: 1 / (i + 1) for i, c in enumerate(chunks)}
    if neighbors:
        chunks += [(d, p + off) for (d, p) in list(chunks) for off in neighbors if (d, p + off) in table]
    unique = sorted(set(chunks), key=lambda c: (c[0], c[1]))
    spans = []
    for _, group in groupby(unique, key=lambda c: c[0]):
        seq = []
        for c in group:
            if not seq or c[1] == seq[-1][1] + 1:
                seq.append(c)
            else:
                spans.append(seq)
                seq = [c]
        spans.append(seq)
    spans.sort(key=lambda span: sum(score.get(c, 0.0) for c in span), reverse=True)
    return [(span, sum(score.get(c, 0.0) for c in span)) for span in spans]
