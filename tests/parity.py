"""Parity checkers: CUDA results vs the float64 oracle, with explicit handling of near-ties."""

from __future__ import annotations

import numpy as np

from oracle import vector_search as ovs

SCORE_TOL = 1e-4   # north_star: scores within 1e-4 (fp32)
TIE_GAP = 5e-6     # below this float64 gap two candidates are interchangeable at the cut


def check_sql_semantics(E, chunk_off, q, got_ids, got_sims, *, k, oversample=4, chunk_max_size=2048,
                        metric="cosine", adapter=None, allowed_chunks=None):
    """Compare one query's result with ``oracle.vector_search_sql`` (float64 adjudication)."""
    ref_ids, ref_sims, ref_rows = ovs.vector_search_sql(
        E, chunk_off, q, num_results=k, oversample=oversample, chunk_max_size=chunk_max_size, metric=metric,
        adapter=adapter, allowed_chunks=allowed_chunks, f64=True)
    q2 = ovs.apply_query_adapter(adapter, q)
    dist = ovs.vector_distances_f64(E, q2, metric)
    num_hits = ovs.num_hits_rule(k, oversample, chunk_max_size)
    rows = np.arange(len(dist))
    if allowed_chunks is not None:
        rows = rows[np.asarray(allowed_chunks, bool)[ovs.row_to_chunk(chunk_off, len(dist))]]
    srt = np.sort(dist[rows])
    vec_gap = (srt[num_hits] - srt[num_hits - 1]) if len(srt) > num_hits else np.inf
    chunk_scores = 1.0 - dist
    got_ids = np.asarray(got_ids)
    got_sims = np.asarray(got_sims, dtype=np.float64)
    # scores of the returned chunks must be their true MaxSim score
    true_of_got = np.array([chunk_scores[chunk_off[c]:chunk_off[c + 1]].max() for c in got_ids]) if len(got_ids) else np.zeros(0)
    if allowed_chunks is not None and len(got_ids):
        assert np.all(np.asarray(allowed_chunks, bool)[got_ids])
    assert np.allclose(got_sims, true_of_got, atol=SCORE_TOL), (got_sims, true_of_got)
    assert np.all(np.diff(got_sims) <= 1e-6), "scores must be descending"
    sim_gaps = np.abs(np.diff(ref_sims)) if len(ref_sims) > 1 else np.array([np.inf])
    if vec_gap > TIE_GAP and (len(sim_gaps) == 0 or sim_gaps.min() > TIE_GAP):
        assert got_ids.tolist() == ref_ids.tolist(), (got_ids, ref_ids)
    else:  # near-ties: same multiset of scores, count may differ by the tied vector
        n = min(len(got_ids), len(ref_ids))
        assert abs(len(got_ids) - len(ref_ids)) <= 1
        assert np.allclose(np.sort(true_of_got)[::-1][:n], np.sort(ref_sims)[::-1][:n], atol=10 * TIE_GAP)
    assert np.allclose(got_sims[: len(ref_sims)], ref_sims[: len(got_sims)], atol=SCORE_TOL)


def check_exact_maxsim(E, chunk_off, q, got_ids, got_sims, *, k, metric="cosine"):
    s = ovs.maxsim_scores(E, chunk_off, q, metric, f64=True)
    ref_ids, ref_s = ovs.maxsim_topk_exact(E, chunk_off, q, k, metric)
    got_ids = np.asarray(got_ids)
    assert len(got_ids) == len(ref_ids)
    assert np.allclose(np.asarray(got_sims, np.float64), s[got_ids], atol=SCORE_TOL)
    srt = np.sort(s)[::-1]
    gap = (srt[k - 1] - srt[k]) if len(srt) > k else np.inf
    inner = np.abs(np.diff(ref_s)).min() if len(ref_s) > 1 else np.inf
    if gap > TIE_GAP:
        assert set(got_ids.tolist()) == set(ref_ids.tolist())
    if gap > TIE_GAP and inner > TIE_GAP:
        assert got_ids.tolist() == ref_ids.tolist()
    assert np.allclose(np.sort(s[got_ids])[::-1], ref_s, atol=10 * TIE_GAP)


def check_sql_from_topn(top_rows, top_dist, row_chunk, got_ids, got_sims, *, k, num_hits):
    """Compare one query's result with the oracle's ``ORDER BY dist LIMIT`` list (``oracle.topn_rows_blocked``
    with ``f32_ties=True``; at least ``num_hits + 1`` rows when the table has that many) through the
    shared ``GROUP BY`` restatement.  Returns True for an exact match (ids, order, count); a mismatch is
    only accepted when the cut or two neighbouring scores tie within ``TIE_GAP`` -- SQL leaves those
    orders unspecified -- and then the score lists must still agree."""
    top_rows, top_dist = np.asarray(top_rows), np.asarray(top_dist)
    chunks = np.asarray(row_chunk)[top_rows] if not callable(row_chunk) else row_chunk(top_rows)
    ref_ids, ref_sims = ovs.group_hits(top_dist[:num_hits], chunks[:num_hits], k)
    got_ids = np.asarray(got_ids)
    got_sims = np.asarray(got_sims, dtype=np.float64)
    assert np.all(np.diff(got_sims) <= 1e-6), "scores must be descending"
    if got_ids.tolist() == ref_ids.tolist():
        assert np.allclose(got_sims, ref_sims, atol=SCORE_TOL), np.abs(got_sims - ref_sims).max()
        return True
    d64 = top_dist.astype(np.float64)
    vec_gap = (d64[num_hits] - d64[num_hits - 1]) if len(d64) > num_hits else np.inf
    sim_gaps = np.abs(np.diff(ref_sims.astype(np.float64))) if len(ref_sims) > 1 else np.array([np.inf])
    assert vec_gap <= TIE_GAP or sim_gaps.min() <= TIE_GAP, ("mismatch without a tie", got_ids[:10], ref_ids[:10], vec_gap)
    # every returned chunk must own one of the listed vectors and carry that vector's similarity
    best = {}
    for c, dd in zip(chunks.tolist(), d64.tolist()):
        best.setdefault(c, 1.0 - dd)
    assert all(int(c) in best for c in got_ids), "a returned chunk owns none of the nearest vectors"
    assert np.allclose(got_sims, [best[int(c)] for c in got_ids], atol=SCORE_TOL)
    n = min(len(got_ids), len(ref_ids))
    if vec_gap > TIE_GAP:   # the vector set is unambiguous, only the order of tied chunks may differ
        assert len(got_ids) == len(ref_ids) and sorted(got_ids.tolist()) == sorted(ref_ids.tolist())
    assert np.allclose(got_sims[:n], ref_sims[:n], atol=SCORE_TOL)
    return False
