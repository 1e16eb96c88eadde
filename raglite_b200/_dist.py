"""Row-sharded corpus across the GPUs of one box: one process per GPU, each scans its shard, then a
SINGLE all-gather of the per-shard hit lists (NCCL over NVLink) and a local merge on every rank.

The reference has no distributed path; this is the B200-native equivalent of
"one big chunk_embedding table": chunks are partitioned into contiguous ranges (never splitting a
chunk's vectors), queries are replicated, and ``GROUP BY chunk / ORDER BY / LIMIT`` (_search.py:143-150)
runs over the gathered top-``num_hits`` vectors, which is exactly what a single table would produce.
"""

from __future__ import annotations

from typing import Any

import torch
import torch.distributed as dist

from ._index import CorpusIndex, ScanResult, limit_hits_to_nearest, merge_hits


def pack_hits(hit_sim: torch.Tensor, hit_chunk: torch.Tensor, hit_count: torch.Tensor) -> torch.Tensor:
    """One contiguous byte buffer per rank: chunk ids (int64) | sims (float32) | counts (int32)."""
    parts = [hit_chunk.contiguous().view(torch.uint8).reshape(-1), hit_sim.contiguous().view(torch.uint8).reshape(-1),
             hit_count.contiguous().view(torch.uint8).reshape(-1)]
    return torch.cat(parts)


def unpack_hits(buf: torch.Tensor, R: int, B: int, H: int) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Inverse of :func:`pack_hits` for ``R`` concatenated rank buffers -> ``[R, B, H]`` views."""
    per = B * H * 12 + B * 4
    buf = buf.reshape(R, per)
    n8, n4 = B * H * 8, B * H * 4
    chunk = buf[:, :n8].contiguous().view(torch.int64).reshape(R, B, H)
    sim = buf[:, n8:n8 + n4].contiguous().view(torch.float32).reshape(R, B, H)
    count = buf[:, n8 + n4:].contiguous().view(torch.int32).reshape(R, B)
    return sim, chunk, count


def gather_hits(hit_sim: torch.Tensor, hit_chunk: torch.Tensor, hit_count: torch.Tensor, group: Any | None
                ) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """The single collective of the path: all-gather every rank's packed hit list."""
    B, H = int(hit_sim.shape[0]), int(hit_sim.shape[1])
    if group is None or dist.get_world_size(group) == 1:
        return hit_sim[None], hit_chunk[None], hit_count[None]
    R = dist.get_world_size(group)
    mine = pack_hits(hit_sim, hit_chunk, hit_count)
    out = torch.empty(R * mine.numel(), dtype=torch.uint8, device=mine.device)
    dist.all_gather_into_tensor(out, mine, group=group)
    return unpack_hits(out, R, B, H)


def shard_ranges(chunk_off: Any, world: int) -> list[tuple[int, int]]:
    """Contiguous chunk ranges with (nearly) equal vector counts; a chunk is never split."""
    import numpy as np

    chunk_off = np.asarray(chunk_off, dtype=np.int64)
    n_chunks = len(chunk_off) - 1
    total = int(chunk_off[-1])
    cuts = [0]
    for r in range(1, world):
        target = total * r // world
        c = int(np.searchsorted(chunk_off, target, side="left"))
        cuts.append(min(max(c, cuts[-1]), n_chunks))
    cuts.append(n_chunks)
    return [(cuts[i], cuts[i + 1]) for i in range(world)]


class ShardedIndex:
    """A ``CorpusIndex`` shard plus the process group it is one part of."""

    def __init__(self, local: CorpusIndex, group: Any | None = None, chunk_ids: list[str] | None = None):
        self.local = local
        self.group = group
        self.world = dist.get_world_size(group) if group is not None else 1
        self.global_chunk_ids = chunk_ids
        self.last_status: torch.Tensor | None = None

    def search_device(self, Q: torch.Tensor, *, k: int, num_hits: int, metric: str = "cosine", algo: str = "auto",
                      row_allowed: torch.Tensor | None = None, checked: bool = True, flags: int = 0,
                      sample_stride: int = 0, rank_first_limit: int | None = None
                      ) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """Scan the local shard, all-gather, merge.  Everything stays on the device / current stream
        (``checked=True`` adds the host-side overflow check and retry).  ``rank_first_limit`` applies the
        rank-then-filter metadata branch to the gathered lists (``limit_hits_to_nearest``)."""
        fn = self.local.scan_checked if checked else self.local.scan
        res: ScanResult = fn(Q, k=k, num_hits=num_hits, metric=metric, algo=algo, row_allowed=row_allowed, flags=flags,
                             sample_stride=sample_stride)
        self.last_status = res.status
        sim, chunk, count = gather_hits(res.hit_sim, res.hit_chunk, res.hit_count, self.group)
        if rank_first_limit is not None:
            count = limit_hits_to_nearest(self, Q, sim, count, k=k, num_hits=num_hits, metric=metric, algo=algo,
                                          limit=rank_first_limit)
        return merge_hits(sim, chunk, count, num_hits=num_hits, k=k)

    def sum_over_shards(self, x: torch.Tensor) -> torch.Tensor:
        """All-reduce (sum) of a small per-query tensor: row counts of the rank-then-filter probe."""
        if self.group is not None and self.world > 1:
            x = x.clone()
            dist.all_reduce(x, op=dist.ReduceOp.SUM, group=self.group)
        return x

    def chunk_id_of(self, global_chunk: int) -> str:
        if self.global_chunk_ids is not None:
            return self.global_chunk_ids[int(global_chunk)]
        lo = self.local.chunk_base
        if self.local.chunk_ids is not None and lo <= global_chunk < lo + self.local.n_chunks:
            return self.local.chunk_ids[int(global_chunk) - lo]
        return str(int(global_chunk))
