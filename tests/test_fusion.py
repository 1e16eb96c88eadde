"""Reciprocal Rank Fusion / hybrid_search and the span collation of retrieve_chunk_spans, batched on
device chunk indices (``rl_rrf_fuse``, ``rl_span_collate``) against the oracle -- the RRF oracle itself is pinned
to outputs of the reference's own function (``tests/golden/rrf.npz``)."""

from __future__ import annotations

import json

import numpy as np
import pytest

from oracle import fusion as ofu


@pytest.fixture(scope="module")
def rrf_cases(golden_dir):
    return json.loads(bytes(np.load(golden_dir / "rrf.npz")["cases"]).decode())


def test_oracle_rrf_is_the_reference_function(rrf_cases):
    for c in rrf_cases:
        ids, scores = ofu.reciprocal_rank_fusion(c["rankings"], k=c["k"], weights=c["weights"])
        assert ids == c["ids"] and scores == c["scores"]          # same floats, bit for bit


def test_oracle_span_collation_example():
    table = {("A", i) for i in range(12)} | {("B", i) for i in range(5)}
    spans = ofu.collate_chunk_spans([("A", 5), ("B", 2), ("A", 6), ("A", 9)], table, neighbors=None)
    assert [s for s, _ in spans] == [[("A", 5), ("A", 6)], [("B", 2)], [("A", 9)]]
    spans = ofu.collate_chunk_spans([("A", 5), ("B", 2), ("A", 9)], table, neighbors=(-1, 1))
    assert [s for s, _ in spans][0] == [("A", 4), ("A", 5), ("A", 6)] and len(spans) == 3


@pytest.mark.gpu
def test_rrf_kernel_matches_the_reference_bit_for_bit(rrf_cases):
    import torch

    import raglite_b200 as rl

    for c in rrf_cases:
        R = len(c["rankings"])
        L = max(1, max(len(r) for r in c["rankings"]))
        t = np.full((1, R, L), -1, np.int64)
        for r, ranking in enumerate(c["rankings"]):
            t[0, r, : len(ranking)] = ranking
        ids, score, count = rl.rrf_fuse_device(torch.from_numpy(t).cuda(), c["weights"], k=c["k"])
        n = int(count[0])
        assert ids[0, :n].tolist() == c["ids"] and score[0, :n].tolist() == c["scores"]
        assert (ids[0, n:] == -1).all()
        got_ids, got_scores = rl.reciprocal_rank_fusion([[str(x) for x in r] for r in c["rankings"]], k=c["k"], weights=c["weights"])
        assert got_ids == [str(x) for x in c["ids"]] and got_scores == c["scores"]
    # a batch: many queries in one launch, with ties (equal weights, disjoint rankings)
    rng = np.random.default_rng(0)
    B, R, L = 64, 2, 40
    t = np.stack([np.stack([rng.permutation(200)[:L] for _ in range(R)]) for _ in range(B)]).astype(np.int64)
    t[:, 1, 30:] = -1
    ids, score, count = rl.rrf_fuse_device(torch.from_numpy(t).cuda(), [0.75, 0.25], num_results=25)
    for b in range(B):
        want_ids, want_scores = ofu.reciprocal_rank_fusion([t[b, 0].tolist(), t[b, 1, :30].tolist()], weights=[0.75, 0.25])
        assert ids[b].tolist() == want_ids[:25] and score[b].tolist() == want_scores[:25] and int(count[b]) == 25


@pytest.mark.gpu
def test_hybrid_search_and_device_span_collation():
    import torch
    from synth import make_corpus, make_queries

    import raglite_b200 as rl

    rng = np.random.default_rng(3)
    E, off = make_corpus(300, (1, 4), 32, seed=5, fp16_round=True)
    C = len(off) - 1
    docs = [f"doc-{c // 9:02d}" for c in range(C)]                      # 9 chunks per document, positions 0..8
    chunks = [rl.Chunk(id=f"c{c}", document_id=docs[c], index=c % 9, body=f"[{c}]") for c in range(C)]
    perm = rng.permutation(C)                                            # the table is not stored in document order
    inv = np.argsort(perm)
    rows = np.concatenate([np.arange(off[c], off[c + 1]) for c in perm])
    off_p = np.concatenate([[0], np.cumsum(np.diff(off)[perm])])
    idx = rl.CorpusIndex(E[rows], off_p, chunk_ids=[chunks[c].id for c in perm], chunks=[chunks[c] for c in perm])
    cfg = rl.RAGLiteConfig(db_url="mem://fusion", reranker=None)
    rl.register_index(cfg, idx)
    q = make_queries(E, 1, seed=6)[0]
    # hybrid_search: vector ranking from the device index, keyword ranking from a registered callable, RRF on the device
    kw = [f"c{c}" for c in rng.permutation(C)[:12]]
    rl.register_keyword_search(cfg, lambda query, *, num_results, metadata_filter=None, config=None: (kw[:num_results], [1.0] * num_results))
    import raglite_b200._search as S

    orig_vs = S.vector_search
    S.vector_search = lambda query, **k2: orig_vs(q, **k2)              # (no text embedder here: route the string to the vector)
    try:
        ids, scores = rl.hybrid_search("what?", num_results=5, config=cfg)
    finally:
        S.vector_search = orig_vs
    vs_ids, _ = rl.vector_search(q, num_results=10, config=cfg)
    want_ids, want_scores = ofu.reciprocal_rank_fusion([vs_ids, kw[:10]], weights=[0.75, 0.25])
    assert ids == want_ids[:5] and scores == want_scores[:5]
    # span collation on the device, through the drop-in and batched
    table = {(docs[c], c % 9) for c in range(C)}
    for trial in range(6):
        picked = [int(c) for c in rng.permutation(C)[: int(rng.integers(1, 14))]]
        nbrs = [(-1, 1), None, (-2, -1, 1), (1,)][trial % 4]
        spans = rl.retrieve_chunk_spans([f"c{c}" for c in picked], neighbors=nbrs, config=cfg)
        want = ofu.collate_chunk_spans([(docs[c], c % 9) for c in picked], table, neighbors=nbrs)
        assert [[(ch.document_id, ch.index) for ch in s.chunks] for s in spans] == [s for s, _ in want]
    ranked = torch.full((4, 10), -1, dtype=torch.int64)
    lists = [[int(inv[c]) for c in rng.permutation(C)[:n]] for n in (10, 3, 7, 1)]   # LOCAL indices of the permuted table
    for b, lst in enumerate(lists):
        ranked[b, : len(lst)] = torch.tensor(lst)
    out = rl.collate_spans_device(idx, ranked.cuda(), neighbors=(-1, 1))
    for b, lst in enumerate(lists):
        want = ofu.collate_chunk_spans([(docs[perm[i]], perm[i] % 9) for i in lst], table, neighbors=(-1, 1))
        ns = int(out["n_span"][b])
        assert ns == len(want)
        member = out["member"][b].tolist()
        for s, (span, score) in enumerate(want):
            st, ln = int(out["span_start"][b, s]), int(out["span_len"][b, s])
            got = [(docs[perm[member[st + j]]], int(perm[member[st + j]]) % 9) for j in range(ln)]
            assert got == span and float(out["span_score"][b, s]) == score
    # a deleted chunk is no neighbour
    idx.delete_chunks(["c4"])
    spans = rl.retrieve_chunk_spans(["c3"], neighbors=(-1, 1), config=cfg)
    assert [[ch.id for ch in s.chunks] for s in spans] == [["c2", "c3"]]
