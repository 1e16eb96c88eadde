"""Two-GPU (NCCL) parity of the sharded path: row-sharded corpus, one all-gather of packed hit lists, merge on
every rank -- against the oracle over the whole corpus.  Skipped on a single-GPU box (the driver's 1-GPU
``-m gpu`` run); ``gpurun --gpus 2 -- python -m pytest tests/test_gpu_dist.py -m gpu`` runs it."""

from __future__ import annotations

import os
import sys
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


def _worker(rank: int, world: int, port: int, tmp: str) -> None:
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    from parity import check_sql_semantics
    from synth import make_corpus, make_queries

    import raglite_b200 as rl
    import raglite_b200._search as S
    from raglite_b200._dist import ShardedIndex, shard_ranges

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    E, off = make_corpus(30_000, (1, 6), 64, seed=3, fp16_round=True)
    C = len(off) - 1
    Q = make_queries(E, 24, seed=4)
    lo, hi = shard_ranges(off, world)[rank]
    r0, r1 = int(off[lo]), int(off[hi])
    tagged = np.arange(C) % 2 == 0
    meta = [{"half": int(tagged[c])} for c in range(lo, hi)]
    local = rl.CorpusIndex(E[r0:r1], off[lo:hi + 1] - r0, chunk_base=lo, chunk_metadata=meta, device=f"cuda:{rank}")
    index = ShardedIndex(local, dist.group.WORLD)
    assert index.ranges == [(a, b - a) for a, b in shard_ranges(off, world)]
    cfg = rl.RAGLiteConfig(reranker=None)
    # 1) plain search, SQL semantics and exact MaxSim, both through the packed all-gather + in-place merge
    for exact in (False, True):
        ids, sims, counts = rl.vector_search_batch(Q, num_results=10, config=cfg, index=index, exact_maxsim=exact)
        if not exact:
            for b in range(len(Q)):
                check_sql_semantics(E, off, Q[b], ids[b, :counts[b]], sims[b, :counts[b]], k=10)
        else:
            from parity import check_exact_maxsim
            for b in range(len(Q)):
                check_exact_maxsim(E, off, Q[b], ids[b, :counts[b]], sims[b, :counts[b]], k=10)
    both = [None] * world
    dist.all_gather_object(both, ids.tolist())
    assert both[0] == both[1], "every rank must hold the same merged result"
    # 1b) three searches in flight on their own streams (one thread: launches and collectives in program order on
    # every rank) return what the serial calls return
    parts = [Q[:8], Q[8:16], Q[16:]]
    want = [rl.vector_search_batch(p, num_results=10, config=cfg, index=index) for p in parts]
    pend = [rl.vector_search_batch_async(torch.from_numpy(p).pin_memory(), num_results=10, config=cfg, index=index) for p in parts]
    for w_, p_ in zip(want, pend, strict=True):
        g_ = p_.result()
        assert np.array_equal(g_[0], w_[0]) and np.array_equal(g_[1], w_[1]) and np.array_equal(g_[2], w_[2])
    # 2) candidate overflow on the shards: the status words travel with the hits, all ranks retry together
    Qd = torch.from_numpy(Q).cuda()
    sim, chunk, count = index.search_device(Qd, k=10, num_hits=40, sample_stride=32, checked=True)
    res = local.scan(Qd, k=10, num_hits=40, sample_stride=32, cand_cap=256)
    assert bool((res.status & 1).any()), "the tiny list must overflow on this shard"
    with local._lock:
        out = index.search_pipeline(Qd, k=10, num_hits=40, sample_stride=32, cand_cap=256)
    assert bool((out[3] & 1).any()) and out[3].shape == (world, len(Q))
    ids, sims, counts = rl.vector_search_batch(Q, num_results=10, config=cfg, index=index)
    for b in range(len(Q)):
        check_sql_semantics(E, off, Q[b], ids[b, :counts[b]], sims[b, :counts[b]], k=10)
    # 3) metadata filter, rank-then-filter branch proven from the all-reduced counters (no probe pass)
    S.FILTER_FIRST_MAX_ROWS, S.RANK_FIRST_LIMIT = 1_000, 40_000
    calls = {"n": 0}
    orig = rl.CorpusIndex.count_at_least

    def counting(self, *a, **k):
        calls["n"] += 1
        return orig(self, *a, **k)

    rl.CorpusIndex.count_at_least = counting
    ids, sims, counts = rl.vector_search_batch(Q, num_results=10, metadata_filter={"half": 1}, config=cfg, index=index)
    assert calls["n"] == 0
    for b in range(len(Q)):
        check_sql_semantics(E, off, Q[b], ids[b, :counts[b]], sims[b, :counts[b]], k=10, allowed_chunks=tagged)
    # ... and the explicit probe when the bound cannot prove it (limit below the bound)
    S.RANK_FIRST_LIMIT = 2_000
    ids, sims, counts = rl.vector_search_batch(Q[:4], num_results=10, metadata_filter={"half": 1}, config=cfg, index=index)
    assert calls["n"] >= 1
    from oracle import vector_search as ovs
    for b in range(4):
        got = ids[b, :counts[b]].tolist()
        opts = [ovs.vector_search_sql(E, off, Q[b], num_results=10, allowed_chunks=tagged, f64=True, filter_first_max=1_000,
                                      rank_first_limit=lim)[0].tolist() for lim in (2_000, 1_999, 2_001, 1_990, 2_010)]
        assert got in opts, (b, got, opts[0])
    dist.barrier()
    dist.destroy_process_group()
    Path(tmp, f"ok{rank}").write_text("ok")


def test_two_gpu_sharded_search(tmp_path):
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    port = 29700 + (os.getpid() % 1000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()
