"""world_size-2 ``gloo`` test (CPU) of the N>1 host logic: shard ranges never split a chunk, the
single all-gather moves packed per-shard hit lists intact, and merging the gathered top-num_hits
vectors reproduces the single-table SQL semantics (what rl_topk_merge computes on the GPU)."""

from __future__ import annotations

import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]


def shard_hits_numpy(E, off, Q, lo, hi, num_hits):
    """Per-shard output of rl_maxsim_topk restated in NumPy: top-num_hits vectors (sim, global chunk)."""
    from oracle import vector_search as ovs

    r0, r1 = int(off[lo]), int(off[hi])
    Es = E[r0:r1]
    r2c = ovs.row_to_chunk(off[lo:hi + 1] - r0, r1 - r0) + lo
    B = len(Q)
    sim = np.full((B, num_hits), -np.inf, np.float32)
    chunk = np.full((B, num_hits), -1, np.int64)
    count = np.zeros(B, np.int32)
    for b, q in enumerate(Q):
        dist_ = ovs.vector_distances_f64(Es, q, "cosine")
        order = np.argsort(dist_, kind="stable")[:num_hits]
        n = len(order)
        sim[b, :n] = (1.0 - dist_[order]).astype(np.float32)
        chunk[b, :n] = r2c[order]
        count[b] = n
    return sim, chunk, count


def merge_numpy(sim, chunk, count, num_hits, k):
    """rl_topk_merge restated: merge R sorted lists, keep num_hits, group by chunk (first = max), top-k."""
    R, B, H = sim.shape
    out = []
    for b in range(B):
        s = np.concatenate([sim[r, b, :count[r, b]] for r in range(R)])
        c = np.concatenate([chunk[r, b, :count[r, b]] for r in range(R)])
        o = np.argsort(-s.astype(np.float64), kind="stable")[:num_hits]
        s, c = s[o], c[o]
        uniq, first = np.unique(c, return_index=True)
        keep = np.sort(first)[:k]
        out.append((c[keep], s[keep]))
    return out


def _worker(rank: int, world: int, port: int, tmp: str) -> None:
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from synth import make_corpus, make_queries

    from oracle import vector_search as ovs
    from raglite_b200._dist import gather_hits, shard_ranges

    E, off = make_corpus(600, (1, 9), 32, seed=3)
    Q = make_queries(E, 5, seed=4)
    k, num_hits = 20, 80
    ranges = shard_ranges(off, world)
    assert ranges[0][0] == 0 and ranges[-1][1] == len(off) - 1
    assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
    lo, hi = ranges[rank]
    sim, chunk, count = shard_hits_numpy(E, off, Q, lo, hi, num_hits)
    g_sim, g_chunk, g_count = gather_hits(torch.from_numpy(sim), torch.from_numpy(chunk), torch.from_numpy(count),
                                          dist.group.WORLD)
    assert g_sim.shape == (world, len(Q), num_hits)
    assert torch.equal(g_sim[rank], torch.from_numpy(sim)) and torch.equal(g_chunk[rank], torch.from_numpy(chunk))
    assert torch.equal(g_count[rank], torch.from_numpy(count))
    merged = merge_numpy(g_sim.numpy(), g_chunk.numpy(), g_count.numpy(), num_hits, k)
    for b, q in enumerate(Q):
        ref_ids, ref_sims, _ = ovs.vector_search_sql(E, off, q, num_results=k, f64=True)
        assert merged[b][0].tolist() == ref_ids.tolist()
        assert np.allclose(merged[b][1], ref_sims, atol=1e-6)
    dist.barrier()
    dist.destroy_process_group()
    Path(tmp, f"ok{rank}").write_text("ok")


def test_two_rank_gather_and_merge(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()


def test_shard_ranges_balance_and_edges():
    sys.path.insert(0, str(ROOT))
    from raglite_b200._dist import shard_ranges

    off = np.concatenate([[0], np.cumsum(np.random.default_rng(0).integers(1, 20, size=1000))])
    for world in (1, 2, 4, 8):
        rng = shard_ranges(off, world)
        rows = [off[b] - off[a] for a, b in rng]
        assert sum(rows) == off[-1] and max(rows) - min(rows) <= 2 * 19
    assert shard_ranges(np.array([0, 5]), 4) == [(0, 0), (0, 0), (0, 0), (0, 1)] or sum(b - a for a, b in shard_ranges(np.array([0, 5]), 4)) == 1


def test_pack_unpack_roundtrip():
    sys.path.insert(0, str(ROOT))
    from raglite_b200._dist import pack_hits, unpack_hits

    g = torch.Generator().manual_seed(0)
    B, H, R = 3, 7, 2
    bufs, want = [], []
    for _ in range(R):
        s = torch.randn((B, H), generator=g)
        c = torch.randint(0, 1 << 40, (B, H), generator=g)
        n = torch.randint(0, H, (B,), generator=g, dtype=torch.int32)
        bufs.append(pack_hits(s, c, n)); want.append((s, c, n))
    sim, chunk, count = unpack_hits(torch.cat(bufs), R, B, H)
    for r in range(R):
        assert torch.equal(sim[r], want[r][0]) and torch.equal(chunk[r], want[r][1]) and torch.equal(count[r], want[r][2])
