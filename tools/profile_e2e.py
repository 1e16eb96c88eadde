"""Where does the end-to-end time of vector_search_batch go beyond the device-timed step?  (1 GPU)"""
import json, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "tests")]
import numpy as np, torch
import raglite_b200 as rl
from synth_torch import gaussian_corpus_torch, queries_near_rows

chunks = int(sys.argv[1]) if len(sys.argv) > 1 else 1_250_000
E = gaussian_corpus_torch(chunks * 12, 1024, seed=0, device="cuda")
idx = rl.CorpusIndex(E, vecs_per_chunk=12)
Q = queries_near_rows(E, 256, seed=1)
Qh = torch.empty((256, 1024), dtype=torch.float32, pin_memory=True); Qh.copy_(Q.cpu())
cfg = rl.RAGLiteConfig(reranker=None)
for _ in range(3):
    rl.vector_search_batch(Qh, num_results=100, config=cfg, index=idx)
torch.cuda.synchronize()
N = 10
t0 = time.perf_counter()
for _ in range(N):
    rl.vector_search_batch(Qh, num_results=100, config=cfg, index=idx)
e2e = (time.perf_counter() - t0) / N * 1e3
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ev[0].record()
for _ in range(N):
    idx.search_pipeline(Q, k=100, num_hits=400)
ev[1].record(); torch.cuda.synchronize()
dev = ev[0].elapsed_time(ev[1]) / N
def phase(fn):
    torch.cuda.synchronize(); t = time.perf_counter(); r = fn(); torch.cuda.synchronize(); return r, (time.perf_counter() - t) * 1e3
ph = {}
Qd, ph["h2d_queries"] = phase(lambda: torch.as_tensor(Qh).to("cuda", non_blocking=True).contiguous())
out, ph["pipeline_launch+run"] = phase(lambda: idx.search_pipeline(Qd, k=100, num_hits=400))
t = time.perf_counter(); out = idx.search_pipeline(Qd, k=100, num_hits=400); ph["pipeline_host_enqueue_only"] = (time.perf_counter() - t) * 1e3
torch.cuda.synchronize()
_, ph["to_host"] = phase(lambda: idx.to_host(*out))
print(json.dumps({"e2e_ms": e2e, "device_ms": dev, "gap_ms": e2e - dev, "phases_ms": ph}))
