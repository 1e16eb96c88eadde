"""Self-consistency of the (unpinned) vector_search oracle: hand cases + structural properties."""

from __future__ import annotations

import numpy as np
import pytest
from synth import make_corpus, make_queries, random_orthogonal

from oracle import vector_search as ovs


def test_num_hits_rule():
    assert ovs.num_hits_rule(3, 4, 2048) == 40      # max(k, 10) floor
    assert ovs.num_hits_rule(20, 4, 2048) == 80
    assert ovs.num_hits_rule(100, 4, 2048) == 400
    assert ovs.num_hits_rule(8, 4, 1024) == 20      # round(2.0) * 10
    assert ovs.num_hits_rule(8, 1, 1024) == 0       # round(0.5) == 0 (banker's rounding)
    assert ovs.num_hits_rule(8, 3, 1024) == 20      # round(1.5) == 2


def test_hand_case_cosine():
    E = np.array([[1, 0], [0, 1], [1, 1], [-1, 0], [0.6, 0.8]], dtype=np.float32)
    off = np.array([0, 2, 3, 5])  # chunk0 = rows 0,1; chunk1 = row 2; chunk2 = rows 3,4
    q = np.array([1.0, 0.0], dtype=np.float32)
    ids, sims, rows = ovs.vector_search_sql(E, off, q, num_results=3)
    assert ids.tolist() == [0, 1, 2]
    assert np.allclose(sims, [1.0, 1 / np.sqrt(2), 0.6], atol=1e-6)
    assert rows.tolist() == [0, 2, 4, 1, 3]
    ids2, s2 = ovs.maxsim_topk_exact(E, off, q, 2)
    assert ids2.tolist() == [0, 1] and np.allclose(s2, [1.0, 1 / np.sqrt(2)])


@pytest.mark.parametrize("metric", ["cosine", "dot", "l2"])
def test_metrics_match_definitions(metric):
    rng = np.random.default_rng(0)
    E = rng.standard_normal((50, 16)).astype(np.float32)
    q = rng.standard_normal(16).astype(np.float32)
    d = ovs.vector_distances(E, q, metric)
    d64 = ovs.vector_distances_f64(E, q, metric)
    assert d.dtype == np.float32 and np.allclose(d, d64, atol=2e-5)
    if metric == "dot":
        assert np.allclose(d64, -(E.astype(np.float64) @ q.astype(np.float64)))
    if metric == "l2":
        assert np.allclose(d64, np.linalg.norm(E.astype(np.float64) - q.astype(np.float64), axis=1))


@pytest.mark.parametrize("vecs", [1, 8, (1, 16)])
@pytest.mark.parametrize("metric", ["cosine", "dot", "l2"])
def test_sql_semantics_is_prefix_of_exact_maxsim(vecs, metric):
    """on an exact scan the SQL result is the first min(k, |S|) entries of the exact
    per-chunk MaxSim ranking."""
    E, off = make_corpus(400, vecs, 32, seed=5)
    Q = make_queries(E, 12, seed=6)
    for q in Q:
        ids, sims, _ = ovs.vector_search_sql(E, off, q, num_results=20, metric=metric, f64=True)
        ex_ids, ex_s = ovs.maxsim_topk_exact(E, off, q, 20, metric)
        assert 1 <= len(ids) <= 20
        assert ids.tolist() == ex_ids[: len(ids)].tolist()
        assert np.allclose(sims, ex_s[: len(ids)], atol=1e-12)
        assert np.all(np.diff(sims) <= 0)


def test_fewer_than_k_results_possible():
    # num_hits = 4 * max(k, 10) = 80 vectors; with 16 vectors per chunk, as few as 5 chunks own them.
    E, off = make_corpus(64, 16, 16, seed=9)
    E[:5 * 16] = E[0] + 1e-3 * E[:5 * 16]           # five chunks hugging one direction
    ids, _, rows = ovs.vector_search_sql(E, off, E[0], num_results=20)
    assert len(rows) == 80 and len(ids) < 20


def test_empty_corpus_and_adapter_dtype():
    ids, sims, _ = ovs.vector_search_sql(np.zeros((0, 8), np.float32), np.array([0]), np.ones(8, np.float32))
    assert len(ids) == 0 and len(sims) == 0     # tests/test_search.py:76-85
    A = random_orthogonal(8)
    q16 = np.linspace(-1, 1, 8).astype(np.float16)
    out = ovs.apply_query_adapter(A, q16)
    assert out.dtype == np.float16              # cast back to the query dtype (_search.py:62)
    assert np.array_equal(out, (A @ q16.astype(np.float64)).astype(np.float16))


def test_metadata_filter_first_branch():
    E, off = make_corpus(100, 4, 16, seed=3)
    q = make_queries(E, 1, seed=4)[0]
    allowed = np.zeros(100, bool)
    allowed[::3] = True
    ids, _, _ = ovs.vector_search_sql(E, off, q, num_results=5, allowed_chunks=allowed)
    assert len(ids) == 5 and all(i % 3 == 0 for i in ids)


def test_blas_batch_matches_loop():
    E, off = make_corpus(300, 8, 32, seed=11)
    Q = make_queries(E, 6, seed=12)
    ids, sc = ovs.blas_batch_topk(E, 8, Q, 10)
    ids_h, sc_h = ovs.blas_batch_topk(E, 8, Q, 10, num_hits=40)
    for b, q in enumerate(Q):
        ex, s = ovs.maxsim_topk_exact(E, off, q, 10)
        assert ids[b].tolist() == ex.tolist() and np.allclose(sc[b], s, atol=1e-5)
        sq, ss, _ = ovs.vector_search_sql(E, off, q, num_results=10)
        n = len(sq)
        assert ids_h[b, :n].tolist() == sq.tolist() and np.allclose(sc_h[b, :n], ss, atol=1e-5)


def test_blocked_oracle_equals_the_unblocked_restatement():
    """``topn_rows_blocked`` (used for corpora that only exist on the device) is the same ORDER BY / LIMIT."""
    for metric in ("cosine", "dot", "l2"):
        E, off = make_corpus(400, (1, 7), 40, seed=3, normalize=(metric == "cosine"))
        Q = make_queries(E, 5, seed=4)
        r2c = ovs.row_to_chunk(off)
        blocks = [(r0, E[r0:r0 + 257]) for r0 in range(0, len(E), 257)]
        for f32 in (False, True):
            top = ovs.topn_rows_blocked(blocks, Q, 80, metric, f32_ties=f32)
            for b, q in enumerate(Q):
                ids, sims, rows = ovs.vector_search_sql(E, off, q, num_results=20, metric=metric, f64=True, f32_ties=f32)
                assert np.array_equal(rows, top[b][0])
                i2, s2 = ovs.group_hits(top[b][1], r2c[top[b][0]], 20)
                assert np.array_equal(ids, i2) and np.allclose(sims, s2, atol=1e-12)
    # a gathered subset of a table: owners given per row
    E, off = make_corpus(100, (1, 4), 16, seed=9)
    q = make_queries(E, 1, seed=10)[0]
    sub = np.sort(np.random.default_rng(0).choice(len(E), size=len(E) // 2, replace=False))
    ids, sims, rows = ovs.vector_search_sql(E[sub], None, q, num_results=5, f64=True, row_chunk=ovs.row_to_chunk(off)[sub])
    allowed_rows = np.zeros(len(E), bool); allowed_rows[sub] = True
    dist = ovs.vector_distances_f64(E, q)
    order = sub[np.argsort(dist[sub], kind="stable")][:40]
    want_ids, want_sims = ovs.group_hits(dist[order], ovs.row_to_chunk(off)[order], 5)
    assert np.array_equal(ids, want_ids) and np.allclose(sims, want_sims)


def test_clustered_generator_has_massive_near_ties():
    from synth import make_clustered_corpus

    E, off, cl = make_clustered_corpus(4000, (1, 9), 64, seed=5, mean_cluster=200, max_cluster=1500)
    assert np.allclose(np.linalg.norm(E, axis=1), 1.0, atol=2e-3) and np.array_equal(E, E.astype(np.float16).astype(np.float32))
    sizes = np.bincount(cl[cl >= 0])
    tight = np.nonzero((np.arange(len(sizes)) % 4 == 0) & (sizes >= 100))[0]
    c = tight[np.argmax(sizes[tight])]
    rows = np.nonzero(cl == c)[0]
    s = E[rows] @ E[rows[0]]
    assert (s > 0.998).all()                      # a tight cluster: every member within 2e-3 of every other
    assert 0.2 < (cl < 0).mean() < 0.3
