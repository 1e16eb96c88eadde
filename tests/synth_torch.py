"""Device-side generators for corpora too large to build on the host in a test (the oracle then reads
the SAME tensor back block by block, so both sides see identical data)."""

from __future__ import annotations

import numpy as np
import torch
from synth import CLUSTER_SPREADS, clustered_plan


def gaussian_corpus_torch(n_rows: int, dim: int, seed: int, device, *, fp16_round: bool = False, step: int = 1 << 19):
    """iid Gaussian rows, unit norm; float32 on ``device``."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    E = torch.empty((n_rows, dim), dtype=torch.float32, device=device)
    for r0 in range(0, n_rows, step):
        blk = torch.randn((min(step, n_rows - r0), dim), generator=g, device=device, dtype=torch.float32)
        blk /= blk.norm(dim=1, keepdim=True)
        E[r0:r0 + blk.shape[0]] = blk.half().float() if fp16_round else blk
    return E


def clustered_corpus_torch(n_rows: int, dim: int, seed: int, device, *, mean_cluster: int = 512, max_cluster: int = 8192,
                           background: float = 0.25, rank: int = 16, fp16_round: bool = True, step: int = 1 << 19,
                           dtype=torch.float32):
    """``synth.make_clustered_corpus`` on the device (same cluster plan, torch noise): tight clusters +
    low-rank background, unit norm, float16-rounded.  Returns ``(E, cluster_of_row numpy)``."""
    cl, n_clusters = clustered_plan(n_rows, seed + 1, mean_cluster=mean_cluster, max_cluster=max_cluster, background=background)
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    centers = torch.randn((max(n_clusters, 1), dim), generator=g, device=device)
    centers /= centers.norm(dim=1, keepdim=True)
    spread = torch.tensor(CLUSTER_SPREADS, device=device)[torch.arange(max(n_clusters, 1), device=device) % len(CLUSTER_SPREADS)]
    basis = torch.randn((rank, dim), generator=g, device=device) / dim ** 0.5
    cl_dev = torch.from_numpy(cl).to(device)
    E = torch.empty((n_rows, dim), dtype=dtype, device=device)
    for r0 in range(0, n_rows, step):
        n = min(step, n_rows - r0)
        c = cl_dev[r0:r0 + n]
        noise = torch.randn((n, dim), generator=g, device=device) / dim ** 0.5
        z = torch.randn((n, rank), generator=g, device=device)
        bg = c < 0
        cc = c.clamp(min=0)
        blk = torch.where(bg[:, None], z @ basis + 0.2 * noise, centers[cc] + spread[cc][:, None] * noise)
        blk /= blk.norm(dim=1, keepdim=True)
        E[r0:r0 + n] = (blk.half().float() if fp16_round else blk).to(dtype)
    return E, cl


def queries_near_rows(E: torch.Tensor, n_queries: int, seed: int, *, noise: float = 0.3, frac_random: float = 0.25,
                      rows: np.ndarray | None = None) -> torch.Tensor:
    """``synth.make_queries`` on the device: queries near random (or given) rows + pure-random ones."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    d = int(E.shape[1])
    Q = torch.randn((n_queries, d), generator=g)
    Q /= Q.norm(dim=1, keepdim=True)
    n_near = n_queries - int(round(frac_random * n_queries))
    if rows is None:
        rows = torch.randint(0, int(E.shape[0]), (n_near,), generator=g).numpy()
    Q = Q.to(E.device)
    if n_near:
        Q[:n_near] = E[torch.from_numpy(np.asarray(rows[:n_near])).to(E.device)].float() + noise * Q[:n_near]
        Q[:n_near] /= Q[:n_near].norm(dim=1, keepdim=True)
    return Q.contiguous()


def host_blocks(E: torch.Tensor, step: int = 1 << 18):
    """Yield ``(first_row, float32 ndarray)`` blocks of a device matrix for the blocked oracle."""
    for r0 in range(0, int(E.shape[0]), step):
        yield r0, E[r0:r0 + step].float().cpu().numpy()
