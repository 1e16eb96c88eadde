"""Candidate-volume diagnostics of the online threshold refinement on synthetic shards."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1])); sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tests"))
import numpy as np, torch
import raglite_b200 as rl

def run(chunks, vecs, dim, B, k, num_hits, S, cap):
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    E = torch.randn((chunks * vecs, dim), generator=g, device="cuda"); E /= E.norm(dim=1, keepdim=True)
    idx = rl.CorpusIndex(E, np.arange(0, chunks * vecs + 1, vecs))
    Q = E[torch.randint(0, E.shape[0], (B,), device="cuda")] + 0.3 * torch.nn.functional.normalize(torch.randn((B, dim), device="cuda"), dim=1)
    Q = torch.nn.functional.normalize(Q, dim=1).contiguous()
    res = idx.scan(Q, k=k, num_hits=num_hits, sample_stride=S, cand_cap=cap)
    torch.cuda.synchronize()
    st = idx.scan_stats()
    print(f"chunks={chunks} d={dim} B={B} K'={num_hits} S={S}: stride={st['sample_stride']} cap={st['cand_cap']} cand_total={st['cand_total']} "
          f"cand_max={st['cand_max']} mean={st['cand_total']/B:.0f} surv={st['survivors_total']/B:.0f} overflow={(res.status.cpu() & 1).sum().item()}")

for S in (0, 32, 128):
    run(100_000, 8, 384, 256, 20, 80, S, 400_000)
run(100_000, 8, 384, 256, 100, 400, 0, 400_000)
run(20_000, 8, 128, 64, 10, 40, 0, 200_000)
run(400_000, 12, 1024, 256, 100, 400, 0, 400_000)
