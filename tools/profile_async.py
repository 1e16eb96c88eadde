"""Where does a pipelined end-to-end step go?  Times N steps of the C4-shard search four ways on one GPU:
device loop on one stream / device loop alternating two streams / async API with device queries / async API with
pinned host queries (2 and 3 in flight) / serial API."""
import sys, time, json
from collections import deque
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import raglite_b200 as rl
from raglite_b200._index import search_async

chunks = int(sys.argv[1]) if len(sys.argv) > 1 else 1_250_000
N = int(sys.argv[2]) if len(sys.argv) > 2 else 20
vecs, d, B, k, H = 12, 1024, 256, 100, 400
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
E = torch.empty((chunks * vecs, d), dtype=torch.float32, device=dev)
for i in range(0, E.shape[0], 1 << 20):
    blk = torch.randn((min(1 << 20, E.shape[0] - i), d), generator=g, device=dev)
    E[i:i + blk.shape[0]] = blk / blk.norm(dim=1, keepdim=True)
off = np.arange(0, E.shape[0] + 1, vecs, dtype=np.int64)
idx = rl.CorpusIndex(E, off, device=dev)
Qd = torch.randn((B, d), generator=g, device=dev); Qd /= Qd.norm(dim=1, keepdim=True)
Qh = torch.empty((B, d), dtype=torch.float32, pin_memory=True); Qh.copy_(Qd.cpu())
cfg = rl.RAGLiteConfig(reranker=None)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

def t(fn, n=N):
    fn(3); torch.cuda.synchronize(); t0 = time.perf_counter(); fn(n); torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3 / n

def dev_one(n):
    for _ in range(n): idx.search_pipeline(Qd, k=k, num_hits=H)
def dev_two(n):
    for i in range(n):
        with torch.cuda.stream(s1 if i % 2 == 0 else s2): idx.search_pipeline(Qd, k=k, num_hits=H)
def mk_async(q, depth):
    def f(n):
        pend = deque()
        for _ in range(n):
            pend.append(search_async(idx, q, k=k, num_hits=H, metric="cosine"))
            if len(pend) >= depth: pend.popleft().result()
        while pend: pend.popleft().result()
    return f
def serial(n):
    for _ in range(n): rl.vector_search_batch(Qh, num_results=k, config=cfg, index=idx)
out = {}
for name, fn in [("device_one_stream", dev_one), ("device_two_streams", dev_two), ("async_devQ_2", mk_async(Qd, 2)),
                 ("async_hostQ_2", mk_async(Qh, 2)), ("async_hostQ_3", mk_async(Qh, 3)), ("serial_api", serial),
                 ("device_one_stream_again", dev_one)]:
    out[name] = round(t(fn), 3)
    print(name, out[name], flush=True)
print(json.dumps(out))
